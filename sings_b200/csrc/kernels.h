// kernels.h -- host-side launch functions of the individual .cu files (internal, C++).
// The public boundary is the C ABI in include/sings_b200.h, implemented in api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgs {

// floats per Gaussian in the geometry record consumed by the blend kernels
constexpr int REC_FLOATS = 16;
// record layout: [0]=pixel x, [1]=pixel y, [2]=-0.5*conic.a, [3]=-conic.b, [4]=-0.5*conic.c,
// [5]=opacity, [6]=pmin (power below which alpha < 1/255 for sure), [7]=r, [8]=g, [9]=b,
// [10]=view depth, [11]=unused, [12]=t = -2 pmin with margin (footprint: d^T conic d <= t),
// [13]=-b/c, [14]=-b/a (edge minimisers of the quadratic form), [15]=flags (bit 0..2: colour
// channel clamped at 0)

// floats per Gaussian in the blend-backward accumulator
constexpr int ACC_FLOATS = 12;
// backward accumulator buffer: P rows of ACC_FLOATS, then one 256-byte line of counters (the item
// ticket of the backward blend) -- zeroed together by one memset per backward
inline size_t acc_rows_bytes(int P) { return ((size_t)(P > 0 ? P : 1) * ACC_FLOATS * 4 + 255) / 256 * 256; }
inline size_t acc_total_bytes(int P) { return acc_rows_bytes(P) + 256; }
// accumulator layout: [0]=dL/dmean2D.x, [1]=.y, [2]=dL/dconic.a, [3]=dL/dconic.b (un-doubled),
// [4]=dL/dconic.c, [5]=dL/dopacity, [6..8]=dL/dcolor rgb, [9..11] unused

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS_N = 8;            // keys per thread in one radix tile: per-Gaussian depth items
constexpr int SORT_ITEMS_L = 16;           // ... and (tile|depth, id) pairs (fewer, larger tiles: shorter look-back)
constexpr int SORT_TILE_N = SORT_ITEMS_N * SORT_THREADS;
constexpr int SORT_TILE_L = SORT_ITEMS_L * SORT_THREADS;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr int DEPTH_PASSES = 4;            // radix passes over the 32 depth bits, run once per Gaussian
// A pair's 32-bit value = Gaussian id (low 24 bits) | reach mask (high 8 bits: which of the tile's
// eight 8x4 pixel blocks the Gaussian's alpha >= 1/255 footprint can touch).  The mask rides
// through the tile-id radix passes for free; the blend kernels read both with one load.
constexpr int ID_BITS = 24;
constexpr unsigned ID_MASK = (1u << ID_BITS) - 1u;

struct RasterLayout {
    // geometry state (per Gaussian)
    size_t rec_off, geom_bytes;
    // zeroed scratch + binning state
    size_t cnt_off, hist_off, scan_off, nstat_off, sortstat_off, bktcnt_off, itemcnt_off, zero_bytes, ranges_off, bktlist_off, itemlist_off;
    size_t blklist_off;      // 8 planes (one per pixel block of a tile) of L_cap (id, list position) entries
    size_t plane_entries;    // entries per plane
    size_t nkeys0_off, nkeys1_off, nvals0_off, nvals1_off, rects_off;   // per-Gaussian depth-sort items
    size_t keys0_off, keys1_off, vals0_off, vals1_off, bin_bytes;
    // image state
    size_t finalT_off, ncontrib_off, nblk_off, img_bytes;
    int scan_blocks, sort_blocks, nsort_blocks, tiles, gx, gy, end_bit, passes;
    // the sorted pair list is in keys1/vals1 when the number of tile-id passes is odd
    bool sorted_in_1() const { return ((passes - DEPTH_PASSES) & 1) != 0; }
};

RasterLayout raster_layout(int P, int W, int H, long long L_cap);
int higher_msb(unsigned n);

struct GeomArgs {
    int P, D, M, W, H;
    float tanfovx, tanfovy, scale_modifier;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* colors_precomp;
    const float* cov3D_precomp;
    const float* view;
    const float* proj;
    const float* campos;
    int prefiltered;
    // 1: shs (forward, backward) were final before the kernel that precedes this one on the
    // stream started, that kernel being one of ours: they may be fetched ahead of the
    // programmatic-dependency wait (common.cuh).  Set through SGS_FLAG_EARLY_PARAMS.
    int early_params = 0;
};

struct LbsFuse;      // kernels_lbs.h: non-null = the deform segment runs inside the geometry kernel
int launch_geometry(const GeomArgs& a, const RasterLayout& lay, long long L_cap, int* radii,
                    char* geom, char* bin, cudaStream_t stream, bool clear = true, const LbsFuse* lf = nullptr);

// depth passes over the P per-Gaussian items (passes whose digit is constant are skipped)
int launch_depth_sort(int P, const RasterLayout& lay, char* bin, cudaStream_t stream, int debug);
// chained scan of tiles touched + (tile|depth, id|reach mask) emission in depth order
int launch_emit_pairs(int P, const RasterLayout& lay, long long L_cap, const char* geom, char* bin,
                      int* host_counters, cudaStream_t stream);
// stable passes over the tile-id digits of the emitted pairs
int launch_tile_sort(const RasterLayout& lay, long long L_cap, const char* geom, char* bin, cudaStream_t stream,
                     int debug);
// stand-alone sort (A/B against CUB): sorts n pairs on key bits [0,end_bit)
int launch_sort_pairs_u64(unsigned long long* keys, unsigned* vals, unsigned long long* keys_tmp,
                          unsigned* vals_tmp, char* scratch, size_t scratch_bytes, long long n,
                          int end_bit, int* result_in_tmp, cudaStream_t stream);
size_t sort_scratch_bytes(long long n);

// tile ranges (+ tiles bucketed by list length)
int launch_tile_ranges(const RasterLayout& lay, long long L_cap, char* bin, cudaStream_t stream);

// for_backward: leave the per-block lists and work items the backward blend walks
int launch_blend_fwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     char* img, const float* bg, float* out_color, float* out_alpha,
                     float* out_depth, bool for_backward, cudaStream_t stream);

int launch_blend_bwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     const char* img, const float* bg, const float* dL_dpix, float* acc,
                     cudaStream_t stream);

struct GeomBwdArgs {
    GeomArgs fwd;
    const int* radii;
    const float* acc;          // (P, ACC_FLOATS) from the blend backward
    float* dL_dmeans3D;        // (P,3)
    float* dL_dmeans2D;        // (P,3)
    float* dL_dcolors;         // (P,3)  (meaningful when colors_precomp was given)
    float* dL_dopacity;        // (P,1)
    float* dL_dcov3D;          // (P,6)
    float* dL_dsh;             // (P,M,3) or null
    float* dL_dscales;         // (P,3)
    float* dL_drots;           // (P,4)
    float* stat_accum;         // (P) xyz_gradient_accum += ||dL_dmeans2D.xy||  | all three or none:
    float* stat_denom;         // (P) denom += 1                                | densification statistics
    float* stat_max_radii;     // (P) max_radii2D = max(., radii)               | of the visible Gaussians
};
int launch_geometry_bwd(const GeomBwdArgs& a, const char* geom, cudaStream_t stream, const LbsFuse* lf = nullptr);

int launch_mark_visible(int P, const float* means3D, const float* view, unsigned char* present,
                        cudaStream_t stream);

// (3,H,W) float image -> (H,W,3) uint8: clamp(0,1) * 255 truncated, optional RGB -> BGR (frame_out.cu)
int launch_frame_to_u8(const float* img, int H, int W, int bgr, unsigned char* out, cudaStream_t stream);

// fused L1 + SSIM image loss (image_loss.cu)
int launch_image_loss_fwd(int H, int W, const float* pred, const void* gt, int gt_u8, const float* mask,
                          const float* bg, float* part, float* gtc, double* sums, float w_l1, float w_ssim,
                          float* loss3, cudaStream_t stream);
int launch_image_loss_bwd(int H, int W, const float* pred, const float* gtc, const float* part,
                          const double* sums, float w_l1, float w_ssim, const float* dloss,
                          float* dL_dpred, float* loss_out, cudaStream_t stream);

// exact K nearest neighbours within a point set + mean neighbour distance (knn.cu)
size_t knn_scratch_bytes(int N, int max_cells);
int knn_grid_resolution(int N, int* max_cells);
int launch_knn(int N, const float* xyz, int K, char* scratch, size_t scratch_bytes, float* mean_dist, int* idx_out,
               float* d2_out, cudaStream_t stream);

// multi-scale tri-plane interpolation (hexplane.cu); aabb, res, plane pointer arrays are HOST memory
int launch_hexplane_fwd(int N, const float* pts, const float* aabb_host, int S, int C, const int* res_host,
                        const float* const* planes_host, float* out, cudaStream_t stream);
int launch_hexplane_bwd(int N, const float* pts, const float* aabb_host, int S, int C, const int* res_host,
                        const float* const* planes_host, const float* d_out, float* const* d_planes_host,
                        float* d_pts, cudaStream_t stream);

// Laplacian terms and L2Norm of the canonical Gaussians (regularizers.cu)
int launch_laplacian_loss_fwd(int n, int C, const int* row_ptr, const int* col_idx, const float* vals,
                              const float* row_w, int mode, const float* x, int ldx, float* y, double* sum,
                              float* loss_out, cudaStream_t stream);
int launch_laplacian_loss_bwd(int n, int C, const int* t_ptr, const int* t_row, const float* t_val,
                              const float* row_w, int mode, const float* y, const float* dloss, float* dx,
                              cudaStream_t stream);
int launch_l2norm_fwd(int N, const float* off, const float* scales, int lds, const float* opacity, float thr_s,
                      float thr_o, float l_off, float l_diff, float l_max, float l_op, double* sums, float* loss_out,
                      cudaStream_t stream);
int launch_l2norm_bwd(int N, const float* off, const float* scales, int lds, int S, const float* opacity, float thr_s,
                      float thr_o, const double* sums, float l_off, float l_diff, float l_max, float l_op,
                      const float* dloss, float* d_off, float* d_scales, float* d_opacity, cudaStream_t stream);

int launch_clear3(void* a, size_t na, void* b, size_t nb, void* c, size_t nc, cudaStream_t stream);
int launch_fold_stats(int P, float* step_accum, float* step_denom, float* step_max_radii, float* accum,
                      float* denom, float* max_radii, cudaStream_t stream);
int launch_densify_stats(int P, const float* grad2d, const int* radii, float* accum, float* denom,
                         float* max_radii, cudaStream_t stream);

}  // namespace sgs

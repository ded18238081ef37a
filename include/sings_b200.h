/*
 * sings_b200.h -- C ABI of libsings_b200.so: the B200 (sm_100a) implementation of SinGS's
 * per-frame avatar hot path (SMPL linear-blend-skinning of every Gaussian, then the 3DGS
 * differentiable tile rasterizer, forward and backward).
 *
 * Boundary rules:
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless it says "host";
 *  - the library never allocates or frees: the caller sizes scratch with sgs_raster_sizes()
 *    / sgs_sort_scratch_bytes() and passes raw pointers (torch tensors' data_ptr());
 *  - every entry point takes the CUDA stream to launch on and never synchronises the host
 *    (except with debug != 0); the only process-wide state is a mutex-guarded cache of kernel
 *    attributes (raised shared-memory limits, occupancy), so host threads may call
 *    concurrently, each on its own stream and scratch;
 *  - return value: 0 = ok, < 0 = argument error (SGS_ERR_*), > 0 = a cudaError_t;
 *    sgs_error_string() decodes both.
 * All floating point is IEEE binary32; matrices are 16 contiguous floats read column-major
 * (which is why SinGS passes transposed torch matrices, datasets/utils.py:37-39).
 *
 * Each entry point cites the reference interface it replaces (paths under /root/reference).
 */
#ifndef SINGS_B200_H
#define SINGS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sgs_stream_t; /* cudaStream_t */

#define SGS_ERR_BAD_ARG -1
#define SGS_ERR_BAD_SH_DEGREE -2
#define SGS_ERR_BAD_JOINTS -3
#define SGS_ERR_MISALIGNED -4
#define SGS_ERR_CAPACITY -5

/* bits of the `debug` argument of sgs_raster_forward / sgs_raster_backward */
#define SGS_FLAG_SYNC_CHECK 1   /* synchronise and check after every stage (the reference's debug=True) */
#define SGS_FLAG_PRECLEARED 2   /* the caller ran sgs_raster_clear on (binning, acc) for this frame: the
                                   entry points skip their own memsets -- a memset between two kernels
                                   costs their overlapped launch, so a per-frame caller clears once, up front */
#define SGS_FLAG_EARLY_PARAMS 4 /* the caller vouches that `shs` was last written before the kernel that
                                   precedes this call on the stream started, and that that kernel is one of
                                   this library's (e.g. sgs_lbs_fwd before the forward, the forward's own
                                   kernels before the backward): the SH rows are then fetched ahead of the
                                   programmatic-dependency wait, while the preceding kernel drains */
#define SGS_FLAG_FORWARD_ONLY 8 /* no backward will follow this forward (inference / animation): the forward
                                   blend then skips the per-block lists and work items it otherwise leaves
                                   for the backward (sgs_raster_backward after such a forward is an error
                                   the library cannot detect: its gradients are garbage) */

int sgs_version(void);
const char* sgs_error_string(int code);

/* Stage timing (measurement only).  A timing handle owns n CUDA events; the rasterizer entry
 * points record them at stage boundaries when given a handle (null = no recording):
 *   forward : [0] start, [1] after the per-Gaussian kernel (preprocess; with the fused entry point
 *             also the deform), [2] after binning (depth passes, pair emission, tile-id
 *             passes), [3] after tile ranges, [4] after the blend;
 *   backward: [5] start, [6] after the blend backward, [7] after the per-Gaussian backward;
 *   fused entry points: [8] frame start (before pose -> A), [10] / [11] around the pose backward.
 * sgs_timing_elapsed_ms waits for event j and returns the time from event i to event j. */
int sgs_timing_create(int n_events, void** handle);
int sgs_timing_destroy(void* handle);
int sgs_timing_record(void* handle, int i, sgs_stream_t stream);
/* Only the events whose bit is set are recorded from now on (default: all).  Every record between
 * two kernels costs their overlapped launch (~4 us), so a coarse measurement -- e.g. one interval
 * spanning several stages -- should enable only its two end points. */
int sgs_timing_set_mask(void* handle, unsigned mask);
int sgs_timing_elapsed_ms(void* handle, int i, int j, float* ms);

/* Record-and-replay of a launch sequence (CUDA graph).  No reference counterpart: the reference
 * launches ~100 eager kernels per frame with host round trips in between
 * (gs_renderer_single.py:87-95 blocks on a device-to-host copy inside the rasterizer; the
 * matrix -> quaternion code of rotations.py:98-149 synchronises on boolean masks).  Everything enqueued on `stream` between
 * sgs_graph_begin and sgs_graph_end -- calls of this library; `stream` must not be the legacy
 * default stream -- becomes one executable graph; sgs_graph_launch replays it on any stream.
 * The entry points of this library are capture-safe: no allocation, no host synchronisation,
 * device-side counts.  Buffers are bound by address. */
int sgs_graph_begin(sgs_stream_t stream);
int sgs_graph_end(sgs_stream_t stream, void** graph_exec);
int sgs_graph_launch(void* graph_exec, sgs_stream_t stream);
int sgs_graph_destroy(void* graph_exec);

/* ---------------------------------------------------------------------------------------
 * Rasterizer.  Replaces the pybind extension `diff_gaussian_rasterization._C`
 * (rasterize_gaussians / rasterize_gaussians_backward / mark_visible) that
 * sings/rec/renderer/gs_renderer_single.py:6-9,84-95 and gs_renderer_multiple.py:6-9,95-121
 * reach through GaussianRasterizer.forward.
 * ------------------------------------------------------------------------------------- */

/* Scratch sizes in bytes for P Gaussians, a W x H image and room for L_cap (tile,Gaussian)
 * pairs: geom (per-Gaussian records), binning (keys/values/ranges/look-back state, and the
 * per-pixel-block lists + work items the forward leaves for the backward: 8 planes of L_cap
 * 8-byte entries, mostly untouched), img (final_T, n_contrib, and the last contributor's index in
 * its block's list), acc (backward accumulator).  Replaces the three resize
 * callbacks of [upstream] rasterize_points.cu (geomBuffer / binningBuffer / imgBuffer). */
int sgs_raster_sizes(int P, int W, int H, long long L_cap, size_t* geom_bytes,
                     size_t* binning_bytes, size_t* img_bytes, size_t* acc_bytes);

/* Byte offsets of the inspectable arrays inside the scratch buffers (for parity tests):
 * info[0]=counters (int32[32] in binning: [0]=num_rendered, [1]=overflow), [1]=keys_unsorted,
 * [2]=vals_unsorted, [3]=keys_sorted, [4]=vals_sorted (point_list), [5]=ranges (uint2[tiles]),
 * [6]=final_T (img), [7]=n_contrib (img), [8]=tiles, [9]=end_bit, [10]=passes,
 * [11]=record floats per Gaussian (geom). */
int sgs_raster_layout_info(int P, int W, int H, long long L_cap, long long* info);

/* Zero what one frame needs zeroed, in ONE kernel launch: the counters / histograms / look-back
 * words at the head of `binning` (for the forward), `acc` (for the backward; null to skip) and
 * an optional caller buffer `extra` of extra_bytes (e.g. the d_A / d_transl accumulators of
 * sgs_lbs_bwd; null to skip); all 16-byte aligned.  Optional: without SGS_FLAG_PRECLEARED the
 * forward and the backward clear their own state. */
int sgs_raster_clear(int P, int W, int H, long long L_cap, void* binning, void* acc, void* extra,
                     size_t extra_bytes, sgs_stream_t stream);

/* Forward: replaces _C.rasterize_gaussians ([upstream] rasterize_points.cu
 * RasterizeGaussiansCUDA -> CudaRasterizer::Rasterizer::forward).  Exactly one of
 * shs (P,M,3) / colors_precomp (P,3) and one of (scales (P,3) + rotations (P,4)) /
 * cov3D_precomp (P,6) is non-null.  out_color (3,H,W) and radii (P) are fully written.
 * out_alpha / out_depth (H,W) are optional (null to skip).  host_counters (host, pinned,
 * 2 ints, optional) receives {num_rendered, overflow} in stream order -- written by the emission
 * kernel itself when the memory is mapped into the device's address space (cudaHostAlloc /
 * torch pin_memory under unified addressing), else by an async copy on `stream`; read it after
 * synchronising.  If overflow != 0 the pair list did not fit L_cap: grow the binning buffer and
 * call again.  `debug`: SGS_FLAG_* bits.  Limits: P < 2^24 (a pair-list entry carries the Gaussian id in
 * 24 bits beside its 8-bit reach mask) and L_cap < 2^30, else SGS_ERR_CAPACITY. */
int sgs_raster_forward(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* campos,
                       float tanfovx, float tanfovy, const float* shs, int prefiltered,
                       long long L_cap, void* geom, void* binning, void* img, float* out_color,
                       int* radii, float* out_alpha, float* out_depth, int* host_counters,
                       sgs_stream_t stream, int debug, void* timing);

/* Backward: replaces _C.rasterize_gaussians_backward ([upstream] Rasterizer::backward).
 * geom/binning/img are the buffers the forward filled; acc is scratch (acc_bytes).  All
 * outputs are fully written: dL_dmeans3D (P,3), dL_dmeans2D (P,3), dL_dcolors (P,3),
 * dL_dopacity (P,1), dL_dcov3D (P,6), dL_dsh (P,M,3) [null when colours were precomputed],
 * dL_dscales (P,3), dL_drots (P,4).
 * xyz_gradient_accum / denom / max_radii2D (P each; all three or all null): when given, the
 * densification statistics of this view (see sgs_densify_stats) are accumulated in the same
 * pass, in place. */
int sgs_raster_backward(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        float tanfovx, float tanfovy, const float* shs, const int* radii,
                        const float* dL_dout_color, long long L_cap, const void* geom,
                        const void* binning, const void* img, void* acc, float* dL_dmeans3D,
                        float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots,
                        float* xyz_gradient_accum, float* denom, float* max_radii2D,
                        sgs_stream_t stream, int debug, void* timing);

/* Replaces _C.mark_visible ([upstream] Rasterizer::markVisible). present: (P) bytes. */
int sgs_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     unsigned char* present, sgs_stream_t stream);

/* Densification statistics of one rendered view, in place.  Replaces
 * sings/rec/models/sings_hybrid.py:1013-1015 (add_densification_stats) and
 * sings/rec/trainer/gs_trainer.py:487-490: for Gaussians with radii > 0,
 * xyz_gradient_accum += ||grad_means2D[:, :2]||, denom += 1, max_radii2D = max(., radii).
 * grad_means2D (P,3); radii (P) int32; the three outputs (P) float32. */
int sgs_densify_stats(int P, const float* grad_means2D, const int* radii,
                      float* xyz_gradient_accum, float* denom, float* max_radii2D,
                      sgs_stream_t stream);

/* Data-parallel step epilogue (views sharded over ranks): after the all-reduce of
 * one step's statistics -- SUM for the xyz_gradient_accum / denom increments
 * (sings_hybrid.py:1013-1015), MAX for the radii (gs_trainer.py:487-490) -- fold them into the
 * persistent accumulators and zero the step buffers, in one launch. */
int sgs_fold_stats(int P, float* step_accum, float* step_denom, float* step_max_radii,
                   float* xyz_gradient_accum, float* denom, float* max_radii2D, sgs_stream_t stream);

/* Animation output, device half: (3,H,W) float image -> (H,W,3) uint8 with the arithmetic of
 * sings/rec/trainer/gs_trainer.py:716-717 -- clamp(0,1), float32 multiply by 255, truncation
 * (`.astype('uint8')`), CHW -> HWC -- and, with bgr != 0, the RGB -> BGR swap of :718
 * (cv2.cvtColor) folded in.  One pass on the device instead of a float32 device->host copy and
 * four numpy passes on the host; `out` (4-byte aligned) is then copied out at a quarter of the bytes. */
int sgs_frame_to_u8(const float* image, int H, int W, int bgr, unsigned char* out, sgs_stream_t stream);

/* ---- multi-scale tri-plane interpolation (first half of the attribute decode).
 * Replaces HexPlaneField.forward of /root/reference/sings/rec/models/modules/hexplane.py:165-189
 * (normalize_aabb :165-166, interpolate_ms_features :70-105 over grid_sample_wrapper :44-68:
 * bilinear, align_corners=True, padding_mode='border'; per scale the PRODUCT of the planes (0,1),
 * (0,2), (1,2); scales concatenated), called over all N Gaussians at sings_hybrid.py:252.
 *   pts (N,3) device; aabb_host: 6 HOST floats = aabb[0] (3), aabb[1] (3) of the module;
 *   res_host: n_scales x 3 HOST ints (reso x, y, z of each scale); C channels per plane (multiple of 32);
 *   planes_host: HOST array of 3 * n_scales DEVICE pointers, scale-major, planes in the order above,
 *   each CHANNEL-LAST (H, W, C) -- i.e. the reference's (1, C, H, W) parameter permuted (0, 2, 3, 1);
 *   out (N, n_scales * C).  1 <= n_scales <= 4.
 * _bwd: d_out (N, n_scales * C) -> d_planes_host (HOST array of device pointers, channel-last like
 * the planes, ACCUMULATED into: zero them first; null = skip) and d_pts (N,3) (null = skip). */
int sgs_hexplane_fwd(int N, const float* pts, const float* aabb_host, int n_scales, int C, const int* res_host,
                     const float* const* planes_host, float* out, sgs_stream_t stream);
int sgs_hexplane_bwd(int N, const float* pts, const float* aabb_host, int n_scales, int C, const int* res_host,
                     const float* const* planes_host, const float* d_out, float* const* d_planes_host, float* d_pts,
                     sgs_stream_t stream);

/* ---- neighbour distances: exact K nearest neighbours of every point of
 * xyz (N,3) among the other points of the set, by Euclidean distance.  Replaces
 * pytorch3d.ops.knn_points(verts[None], verts[None], K + 1) at
 * /root/reference/sings/rec/losses/loss_items.py:75 (whose first column is the point itself; K here
 * counts the real neighbours: the reference's K=9 is K=8) and :78-79:
 *   mean_dist (N)   mean over the K neighbours of |x_j - x_i|      (nullable)
 *   idx (N,K) int   neighbours, nearest first (-1 where the set has fewer than K + 1 points; ties in
 *                   arbitrary order)                                (nullable)
 *   dist2 (N,K)     their squared distances, ascending              (nullable)
 * With fewer than K + 1 points the missing neighbours have idx -1 and dist2 3.4e38 (and mean_dist is
 * meaningless): callers need N > K.  1 <= K <= 16.  scratch: sgs_knn_scratch_bytes(N) bytes, 256-byte aligned. */
size_t sgs_knn_scratch_bytes(int N);
int sgs_knn_mean_dist(int N, const float* xyz, int K, void* scratch, size_t scratch_bytes, float* mean_dist,
                      int* idx, float* dist2, sgs_stream_t stream);

/* ---- image loss: fused L1 + SSIM of the rendered image against the masked
 * ground truth.  Replaces /root/reference/sings/rec/losses/loss.py:57-70 (HumanLoss.forward:
 * gt' = gt m + bg (1 - m); l1 = sum |pred - gt'| / sum m; ssim term = (1 - mean ssim_map) *
 * (sum m / (H W))) with l1_loss / ssim of /root/reference/sings/rec/losses/utils.py:16-70 (11x11
 * window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2).
 *   pred (3,H,W) float; gt (3,H,W) float, or (H,W,3) uint8 when gt_is_u8_hwc (value / 255);
 *   mask (H,W) float or null (= ones); bg (3); scratch: sgs_image_loss_scratch_floats(H, W) floats,
 *   written by _fwd and read by _bwd; sums: 4 doubles, 8-byte aligned, written by _fwd:
 *   [0] = sum |pred - gt'|, [1] = sum of the SSIM map over 3 H W values, [2] = sum m.
 * The loss is  w_l1 sums[0] / sums[2] + w_ssim (1 - sums[1] / (3 H W)) (sums[2] / (H W));  if loss3
 * (3 device floats) is given, _fwd also writes { loss, l1 = sums[0] / sums[2], ssim term } there
 * (everything stays on the device).  _bwd writes dL/dpred (3,H,W) of
 * that loss times *dloss (device scalar, null = 1) and, if loss_out is given, the loss itself. */
size_t sgs_image_loss_scratch_floats(int H, int W);
int sgs_image_loss_fwd(int H, int W, const float* pred, const void* gt, int gt_is_u8_hwc, const float* mask,
                       const float* bg, float* scratch, double* sums, float w_l1, float w_ssim, float* loss3,
                       sgs_stream_t stream);
int sgs_image_loss_bwd(int H, int W, const float* pred, const float* scratch, const double* sums,
                       float w_l1, float w_ssim, const float* dloss, float* dL_dpred, float* loss_out,
                       sgs_stream_t stream);

/* ---- regularisers on the canonical Gaussians (the non-image terms of
 * gs_trainer.py:363-396).
 *
 * Laplacian terms: y = L x for a FIXED sparse operator L (n x n, CSR: row_ptr (n+1), col_idx, vals;
 * built once per densification by the host from pytorch3d.ops.laplacian's definition L = D^-1 A - I),
 * loss = sum_r row_w[r] f(y_r):
 *   mode 0  f = |y_r|^2 : RegionLaplacianLoss_v2.forward / forward_hands,
 *           /root/reference/sings/rec/losses/loss_items.py:173-190 -- sum over regions of
 *           w_region * mean((L_region x_region)^2); all regions are ONE operator with
 *           row_w[r] = w_region(r) / (n_region(r) * C)
 *   mode 1  f = |y_r|   : pcd_laplacian_smoothing, loss_items.py:205-214 (row_w = 1 / n)
 *   x (n, C) floats with row stride ldx >= C (so shs[:n, 0] is taken in place), 1 <= C <= 4;
 *   y (n, C) dense, written by _fwd and read by _bwd; sum: 1 double (8-byte aligned) = the loss;
 *   loss_out (1 device float, nullable) = (float)sum.
 * _bwd: dx (n, C) dense = *dloss (device scalar, null = 1) * L^T (row_w f'(y)) -- a gather over
 * the CSR of L^T (t_ptr (n+1), t_row, t_val), no atomics; f'(0) = 0 in mode 1 (torch's norm backward). */
int sgs_laplacian_loss_fwd(int n, int C, const int* row_ptr, const int* col_idx, const float* vals,
                           const float* row_w, int mode, const float* x, int ldx, float* y, double* sum,
                           float* loss_out, sgs_stream_t stream);
int sgs_laplacian_loss_bwd(int n, int C, const int* t_ptr, const int* t_row, const float* t_val,
                           const float* row_w, int mode, const float* y, const float* dloss, float* dx,
                           sgs_stream_t stream);

/* L2Norm.forward, /root/reference/sings/rec/losses/loss_items.py:36-54, with s = scales[:, 0]:
 *   loss = lambda_xyz_offsets |xyz_offsets| + lambda_scales_diff |s - mean(s)|
 *        + lambda_max_scale |s[s > max_scale_threshold]| + lambda_min_opacity |0.5 - o[o < min_opacity_threshold]|
 * (Frobenius norms).  xyz_offsets (N,3) dense; scales: column 0 is read with row stride lds floats;
 * opacity (N); each of the three may be null (its terms are 0, as when the key is absent from the
 * reference's dict).  sums: 9 doubles (8-byte aligned) written by _fwd and read by _bwd -- [0..4] the
 * five sums, [5..8] the four norms; loss_out (1 device float, nullable).
 * _bwd: d_xyz_offsets (N,3), d_scales (N, scale_cols; column 0 carries the gradient, the others are
 * zeroed), d_opacity (N), each nullable, times *dloss (device scalar, null = 1); the gradient of a
 * norm that is 0 is 0 (torch's norm backward). */
int sgs_l2norm_fwd(int N, const float* xyz_offsets, const float* scales, int lds, const float* opacity,
                   float max_scale_threshold, float min_opacity_threshold, float lambda_xyz_offsets,
                   float lambda_scales_diff, float lambda_max_scale, float lambda_min_opacity, double* sums,
                   float* loss_out, sgs_stream_t stream);
int sgs_l2norm_bwd(int N, const float* xyz_offsets, const float* scales, int lds, int scale_cols,
                   const float* opacity, float max_scale_threshold, float min_opacity_threshold,
                   const double* sums, float lambda_xyz_offsets, float lambda_scales_diff, float lambda_max_scale,
                   float lambda_min_opacity, const float* dloss, float* d_xyz_offsets, float* d_scales,
                   float* d_opacity, sgs_stream_t stream);

/* Stand-alone stable radix sort of n (u64 key, u32 value) pairs on key bits [0,end_bit):
 * what the rasterizer uses in place of cub::DeviceRadixSort::SortPairs ([upstream]
 * rasterizer_impl.cu).  The result is in (keys,vals) when *result_in_tmp (host) == 0, else
 * in (keys_tmp, vals_tmp). */
size_t sgs_sort_scratch_bytes(long long n);
int sgs_sort_pairs_u64(unsigned long long* keys, unsigned int* vals,
                       unsigned long long* keys_tmp, unsigned int* vals_tmp, void* scratch,
                       size_t scratch_bytes, long long n, int end_bit, int* result_in_tmp,
                       sgs_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Deformer.  Replaces the torch ops of sings/rec/models/sings_hybrid.py:398-428 (forward)
 * and :525-552 (forward_chunk): lbs_extra (sings/rec/utils/body_model/lbs.py:16-74),
 * matrix_to_quaternion / quaternion_multiply (sings/rec/utils/geometry/rotations.py:98-149,
 * 393-407), and the pose -> A chain batch_rodrigues + batch_rigid_transform
 * (sings/rec/utils/body_model/smpl.py:415-446,462-513).
 * ------------------------------------------------------------------------------------- */

/* pose (B,J,3) axis-angle, rest (J,3) rest joints, parents (J) int32 (parents[0] = -1,
 * parents[j] < j), inv_A_t2cano (J,16) or null -> A_out (B,J,16) row-major 4x4
 * = A_t2pose @ inv_A_t2cano; G_out (B,J,12) optional (needed by the backward). */
int sgs_pose_to_A(const float* pose, const float* rest, const int* parents,
                  const float* inv_A_t2cano, int B, int J, float* A_out, float* G_out,
                  sgs_stream_t stream);
int sgs_pose_to_A_bwd(const float* pose, const float* rest, const int* parents,
                      const float* inv_A_t2cano, const float* G, const float* dL_dA, int B, int J,
                      float* dL_dpose, sgs_stream_t stream);

/* Fused LBS forward for B frames.  A (B,J,16); xyz_canon (N,3); W (N,J); rot_canon (N,9) or
 * null (identity, the isotropic case); scales (N,3); smpl_scale (B) or null; transl (B,3) or
 * null; ext_* (trans (B,3), rot (B,9), scale (B)) all null or all set.
 * Outputs xyz (B,N,3), rotq (B,N,4), scales_out (B,N,3), T_out (B,N,16) or null. */
int sgs_lbs_fwd(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                const float* rot_canon, const float* scales, const float* smpl_scale,
                const float* transl, const float* ext_trans, const float* ext_rot,
                const float* ext_scale, float* xyz_out, float* rotq_out, float* scales_out,
                float* T_out, sgs_stream_t stream);

/* sgs_pose_to_A immediately followed by sgs_lbs_fwd on A_out (the per-frame deform segment,
 * sings_hybrid.py:398-428, in one call).  Same results as the two calls; because the library
 * knows the kernel that precedes the LBS kernel, the LBS kernel fetches its canonical-parameter
 * tiles while pose -> A is still running (programmatic dependent launch).  The canonical
 * arrays must not be written by work enqueued on `stream` after the call preceding this one. */
int sgs_pose_lbs_fwd(const float* pose, const float* rest, const int* parents,
                     const float* inv_A_t2cano, int B, int N, int J, float* A_out, float* G_out,
                     const float* xyz_canon, const float* W, const float* rot_canon,
                     const float* scales, const float* smpl_scale, const float* transl,
                     float* xyz_out, float* rotq_out, float* scales_out, sgs_stream_t stream);

/* Backward of sgs_lbs_fwd for upstream gradients g_xyz, g_rotq, g_scales and (optional,
 * null if T was not used) g_T (B,N,16).  Written:
 * d_xyz_canon (N,3), d_rot_canon (N,9) or null, d_scales (N,3).  Accumulated into
 * (caller zero-fills): d_A (B,J,16), d_smpl_scale (B) or null, d_transl (B,3) or null. */
int sgs_lbs_bwd(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                const float* rot_canon, const float* scales, const float* smpl_scale,
                const float* transl, const float* ext_trans, const float* ext_rot,
                const float* ext_scale, const float* g_xyz, const float* g_rotq,
                const float* g_scales, const float* g_T, float* d_xyz_canon, float* d_rot_canon,
                float* d_scales, float* d_A, float* d_smpl_scale, float* d_transl,
                sgs_stream_t stream);

/* The same two calls with the canonical rotation given in the 6D representation the model
 * stores (rot6d_canon (N,6); sings_hybrid.py:354-356: rotation_6d_to_matrix of
 * sings/rec/utils/geometry/rotations.py:545-566, Gram-Schmidt with rows b1, b2, b3) -- the
 * (N,3,3) matrix is never materialised.  d_rot6d_canon is (N,6). */
int sgs_lbs_fwd_rot6d(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                      const float* rot6d_canon, const float* scales, const float* smpl_scale,
                      const float* transl, const float* ext_trans, const float* ext_rot,
                      const float* ext_scale, float* xyz_out, float* rotq_out, float* scales_out,
                      float* T_out, sgs_stream_t stream);
int sgs_lbs_bwd_rot6d(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                      const float* rot6d_canon, const float* scales, const float* smpl_scale,
                      const float* transl, const float* ext_trans, const float* ext_rot,
                      const float* ext_scale, const float* g_xyz, const float* g_rotq,
                      const float* g_scales, const float* g_T, float* d_xyz_canon,
                      float* d_rot6d_canon, float* d_scales, float* d_A, float* d_smpl_scale,
                      float* d_transl, sgs_stream_t stream);

/* 6D rotation conversions of sings/rec/utils/geometry/rotations.py, n rotations each:
 * rotation_6d_to_matrix (:545-566) d6 (n,6) -> R (n,9) row-major, and
 * rotation_6d_to_axis_angle (:601-603 = matrix_to_quaternion :98-149 + quaternion_to_axis_angle
 * :514-542) d6 (n,6) -> aa (n,3): the per-frame conversion of the stored global_orient /
 * body_pose parameters (sings_hybrid.py:370-376).  The _bwd calls write dL/dd6 (n,6). */
int sgs_rot6d_to_matrix(const float* d6, int n, float* R_out, sgs_stream_t stream);
int sgs_rot6d_to_matrix_bwd(const float* d6, const float* dL_dR, int n, float* dL_dd6,
                            sgs_stream_t stream);
int sgs_rot6d_to_axis_angle(const float* d6, int n, float* aa_out, sgs_stream_t stream);
int sgs_rot6d_to_axis_angle_bwd(const float* d6, const float* dL_daa, int n, float* dL_dd6,
                                sgs_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * The whole frame in two calls: pose -> A, the deform segment and the rasterizer with the
 * deformer FUSED into the rasterizer's per-Gaussian kernels (forward: LBS in the prologue of the
 * preprocess kernel; backward: the LBS backward in the epilogue of the preprocess backward), so
 * the deformed means / quaternions / scales and their gradients make no round trip through
 * memory between two kernels.  Same results as sgs_pose_lbs_fwd + sgs_raster_forward and
 * sgs_raster_backward + sgs_lbs_bwd + sgs_pose_to_A_bwd (one iteration of the reference's hot
 * loop, gs_trainer.py:229-244 and :400; sings_hybrid.py:398-419; no ext_tfs).
 *
 * The skinning weights come packed: lbs_weights is constant between densifications
 * (sings_hybrid.py:724) with a handful of non-zero entries per row, so the caller packs it once
 * with sgs_lbs_pack_weights into K slots per Gaussian (K in {4, 8, 12, 16}); *max_nnz (device
 * int, zeroed by the caller) receives the longest row -- pack again with a larger K, or stay on
 * the unfused entry points, if it exceeds K.  wq and iq take sgs_lbs_packed_bytes(N, K) and
 * sgs_lbs_packed_bytes(N, K) / 4 bytes.
 * ------------------------------------------------------------------------------------- */
size_t sgs_lbs_packed_bytes(int N, int K);
int sgs_lbs_pack_weights(int N, int J, const float* W, int K, float* wq, unsigned int* iq, int* max_nnz,
                         sgs_stream_t stream);

typedef struct sgs_deform_args {
    int N, J, K, rot6d;          /* Gaussians, joints, packed slots, rot_canon is (N,6) */
    const float* pose;           /* (J,3) axis-angle, joint 0 = global orientation */
    const float* rest;           /* (J,3) rest joints */
    const int* parents;          /* (J) */
    const float* inv_A_t2cano;   /* (J,16) or null */
    const float* xyz_canon;      /* (N,3) */
    const float* scales;         /* (N,3) */
    const float* rot_canon;      /* (N,9) | (N,6) | null (identity) */
    const float* wq;             /* packed weights */
    const unsigned int* iq;      /* packed joint indices */
    const float* smpl_scale;     /* (1) or null */
    const float* transl;         /* (3) or null */
    float* A;                    /* (J,16) out: cano->pose transforms (read again by the backward) */
    float* G;                    /* (J,12) out: global joint transforms (read again by the backward) */
    float* xyz;                  /* (N,3) out: deformed means = the rasterizer's means3D */
    float* rotq;                 /* (N,4) out */
    float* scales_out;           /* (N,3) out */
    float* d_xyz_canon;          /* backward outputs: (N,3) */
    float* d_rot_canon;          /* (N,9) | (N,6) | null */
    float* d_scales;             /* (N,3) */
    float* d_A;                  /* (J,16) accumulated: the caller zeroes it (sgs_raster_clear `extra`) */
    float* d_transl;             /* (3) accumulated likewise, or null */
    float* d_pose;               /* (J,3) */
} sgs_deform_args;

/* Rasterizer arguments as in sgs_raster_forward / sgs_raster_backward (SH colours, scales +
 * quaternions; stage events: [8] frame start, [10] before and [11] after the pose backward). */
int sgs_avatar_forward(const sgs_deform_args* d, int D, int M, int W, int H, const float* bg,
                       const float* opacities, float scale_modifier, const float* viewmatrix,
                       const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                       const float* shs, long long L_cap, void* geom, void* binning, void* img,
                       float* out_color, int* radii, float* out_alpha, float* out_depth, int* host_counters,
                       sgs_stream_t stream, int debug, void* timing);
int sgs_avatar_backward(const sgs_deform_args* d, int D, int M, int W, int H, const float* bg,
                        float scale_modifier, const float* viewmatrix, const float* projmatrix,
                        const float* campos, float tanfovx, float tanfovy, const float* shs, const int* radii,
                        const float* dL_dout_color, long long L_cap, const void* geom, const void* binning,
                        const void* img, void* acc, float* dL_dmeans2D, float* dL_dopacity, float* dL_dsh,
                        float* xyz_gradient_accum, float* denom, float* max_radii2D, sgs_stream_t stream,
                        int debug, void* timing);

#ifdef __cplusplus
}
#endif
#endif /* SINGS_B200_H */

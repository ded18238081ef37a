// raster_geometry.cu -- forward "geometry" stage of the rasterizer, two kernels:
//   geometry_kernel: preprocess (cull, 3D cov, EWA 2D cov, conic, radius, tile rect, SH->RGB)
//     + one (depth, Gaussian id) sort item and one packed tile rectangle per Gaussian
//     + the digit histograms of the four depth passes -- ONE pass over HBM;
//   emit_pairs_kernel (after the Gaussians have been sorted by depth, radix_sort.cu):
//     tiles-touched prefix sum (single-pass chained scan, decoupled look-back)
//     + (tile|depth) key / Gaussian-id emission for every touched tile, in depth order,
//     + the digit histograms of the tile-id passes.
// Together with the radix passes they replace, with identical results, the reference's
//   [upstream] forward.cu preprocessCUDA, cub::DeviceScan::InclusiveSum,
//   rasterizer_impl.cu duplicateWithKeys + cub::DeviceRadixSort::SortPairs
//   (K3..K6 of section 2.4),
// reached from /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// Why the order of work differs from upstream: all pairs of a Gaussian share its depth, so the
// four depth digits of the 64-bit (tile|depth) key are sorted ONCE PER GAUSSIAN (N items)
// instead of once per pair (L ~ 5.5 N items); pairs emitted in depth order then need only the
// stable passes over the tile-id digits.  A stable LSD sort has a unique answer: the sorted
// key/value lists are bit-identical to sorting the pairs on all 45 bits.
//
// Layout: one CTA = 256 consecutive Gaussians.  SH rows (192 B each at M=16) are staged with
// 16-byte cp.async into padded shared rows; the other attributes are read directly.
// Emission CTAs are taken in ticket order so the chained scan can never wait on a CTA that
// has not started, and emission is load-balanced over the CTA (binary search of the
// block-local offsets), so key and value stores are fully coalesced.
#include "geom_math.cuh"
#include "kernels.h"

namespace sgs {

constexpr int GEO_THREADS = 256;
// A/B knobs (tools/sweep.sh)
#ifndef SGS_EMIT_WIDE             // emission scan: look-back by the whole CTA (1) or one warp (0)
#define SGS_EMIT_WIDE 0
#endif
#ifndef SGS_EMIT_MATCH            // tile-digit histogram: one shared atomic per group of equal digits
#define SGS_EMIT_MATCH 0
#endif
// Pair list by count / scan / scatter (1) or by emit_pairs + tile-id radix passes (0).  Both give
// the same bytes (all GPU parity tests pass either way).  Measured at 200k Gaussians / 1024^2
// (profiles/README.md, v7): count 26 us + scan 8.5 us + scatter 52 us = 86 us against 53 us for
// emit + two passes -- the enumeration is ~6k dependent instructions per warp (binary search,
// division, match_any, counter update) with only 8-16 warps per SM because of the 64-100 KB
// counter arrays, i.e. latency-bound at 19 % issue.  Default off until that chain is shortened.
#ifndef SGS_BIN_CSS
#define SGS_BIN_CSS 0
#endif
constexpr int BIN_THREADS = 256;
constexpr int BIN_GAUSS = 1024;           // depth-ordered Gaussians per chunk (CTA): 128 per warp
constexpr int BIN_MAX_TILES = 8192;       // 8 warps x tiles u16 counters + tiles u32 bases must fit in shared memory
constexpr unsigned long long FLAG_AGG = 1ull << 62;
constexpr unsigned long long FLAG_INCL = 2ull << 62;
constexpr unsigned long long FLAG_MASK = 3ull << 62;

int higher_msb(unsigned n) {
    unsigned msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

RasterLayout raster_layout(int P, int W, int H, long long L_cap) {
    RasterLayout l{};
    l.gx = (W + TILE - 1) / TILE;
    l.gy = (H + TILE - 1) / TILE;
    l.tiles = l.gx * l.gy;
    l.end_bit = 32 + higher_msb((unsigned)l.tiles);
    l.passes = (l.end_bit + RADIX_BITS - 1) / RADIX_BITS;
    l.scan_blocks = (P + GEO_THREADS - 1) / GEO_THREADS;
    if (l.scan_blocks < 1) l.scan_blocks = 1;
    l.sort_blocks = (int)((L_cap + SORT_TILE_L - 1) / SORT_TILE_L);
    if (l.sort_blocks < 1) l.sort_blocks = 1;
    l.nsort_blocks = (P + SORT_TILE_N - 1) / SORT_TILE_N;
    if (l.nsort_blocks < 1) l.nsort_blocks = 1;
    // geometry state
    l.rec_off = 0;
    l.geom_bytes = align_up((size_t)(P > 0 ? P : 1) * REC_FLOATS * 4, 256);
    // binning state: zeroed region first
    size_t o = 0;
    l.cnt_off = o;      o += align_up(CNT_SLOTS * 4, 256);
    l.hist_off = o;     o += align_up((size_t)MAX_PASSES * RADIX * 4, 256);
    l.scan_off = o;     o += align_up((size_t)l.scan_blocks * 8, 256);
    l.nstat_off = o;    o += align_up((size_t)DEPTH_PASSES * l.nsort_blocks * RADIX * 4, 256);
    l.sortstat_off = o; o += align_up((size_t)(l.passes - DEPTH_PASSES) * l.sort_blocks * RADIX * 4, 256);
    l.bktcnt_off = o;   o += align_up(32 * 4, 256);
    l.zero_bytes = o;
    size_t np = (size_t)(P > 0 ? P : 1);
    l.nkeys0_off = o;   o += align_up(np * 4, 256);
    l.nkeys1_off = o;   o += align_up(np * 4, 256);
    l.nvals0_off = o;   o += align_up(np * 4, 256);
    l.nvals1_off = o;   o += align_up(np * 4, 256);
    l.rects_off = o;    o += align_up(np * 8, 256);
    l.ranges_off = o;   o += align_up((size_t)l.tiles * 8, 256);
    l.bktlist_off = o;  o += align_up((size_t)32 * l.tiles * 4, 256);
    size_t cap = (size_t)(L_cap > 0 ? L_cap : 1);
    l.keys0_off = o;    o += align_up(cap * 8, 256);
    l.keys1_off = o;    o += align_up(cap * 8, 256);
    l.vals0_off = o;    o += align_up(cap * 4, 256);
    l.vals1_off = o;    o += align_up(cap * 4, 256);
    l.masks_off = o;    o += align_up(cap, 256);
    l.bin_ctas = (P + BIN_GAUSS - 1) / BIN_GAUSS;
    if (l.bin_ctas < 1) l.bin_ctas = 1;
    l.bcount_off = o;   o += align_up((size_t)l.bin_ctas * l.tiles * 2, 256);
    l.bbase_off = o;    o += align_up((size_t)l.bin_ctas * l.tiles * 4, 256);
    l.btotal_off = o;   o += align_up((size_t)l.tiles * 4, 256);
    l.bin_bytes = o;
    // image state
    size_t pix = (size_t)W * H;
    l.finalT_off = 0;
    l.ncontrib_off = align_up(pix * 4, 256);
    l.img_bytes = l.ncontrib_off + align_up(pix * 4, 256);
    return l;
}

struct GeoOut {
    int* radii;
    float* rec;
    unsigned* hist;          // digit histograms, [pass][256]; this kernel fills the depth passes
    unsigned* varbits;       // [0] = OR of the visible depth keys, [1] = OR of their complements
    unsigned* nkeys;         // (P) depth bits (0 for a Gaussian that touches no tile)
    unsigned* nvals;         // (P) Gaussian id
    uint2* rects;            // (P) x0 | y0 << 16, width | height << 16 (tiles)
    int gx, gy;
    float fx, fy;
};

// NVEC: float4 per SH row staged to shared memory; 0 = colours precomputed.
// VEC16: SH rows are 16-byte aligned (M*3 % 4 == 0 and base aligned) -> 16-byte cp.async.
template <int D, bool HAS_SH, bool VEC16>
__global__ void __launch_bounds__(GEO_THREADS)
geometry_kernel(GeomArgs a, GeoOut o) {
    constexpr int NB = (D + 1) * (D + 1);
    constexpr int NVEC = HAS_SH ? sh_nvec(D) : 0;
    constexpr int S4 = HAS_SH ? sh_stride4(NVEC) : 0;
    extern __shared__ float4 s_sh[];                 // GEO_THREADS * S4 float4
    __shared__ float s_cam[36];
    __shared__ unsigned s_hist[DEPTH_PASSES * RADIX];
    __shared__ unsigned s_var[2];

    const int tid = threadIdx.x;
    const int base = blockIdx.x * GEO_THREADS;
    // ---- stage this CTA's SH rows (asynchronously; consumed after the geometry math).  With
    // early_params (the caller vouches that shs was final two library kernels ago, common.cuh)
    // the copies are issued AHEAD of the dependency wait: the largest operand of the kernel is
    // in flight while the preceding kernel (the LBS forward) drains.
    auto stage_sh = [&]() {
        if constexpr (HAS_SH) {
            const int rows = min(GEO_THREADS, a.P - base);
            const size_t row_floats = (size_t)a.M * 3;
            if (VEC16) {
                const int total = rows * NVEC;
                for (int f = tid; f < total; f += GEO_THREADS) {
                    int row = f / NVEC, col = f - row * NVEC;
                    cp_async16(&s_sh[row * S4 + col], a.shs + (size_t)(base + row) * row_floats + col * 4);
                }
            } else {
                float* s_f = reinterpret_cast<float*>(s_sh);
                const int total = rows * NB * 3;
                for (int f = tid; f < total; f += GEO_THREADS) {
                    int row = f / (NB * 3), col = f - row * (NB * 3);
                    cp_async4(&s_f[row * S4 * 4 + col], a.shs + (size_t)(base + row) * row_floats + col);
                }
            }
            cp_async_commit();
        }
    };
    if (a.early_params) stage_sh();
    pdl_sync();
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    else if (tid < 37) s_var[tid - 35] = 0;
    for (int i = tid; i < DEPTH_PASSES * RADIX; i += GEO_THREADS) s_hist[i] = 0;
    __syncthreads();
    const int idx = base + tid;
    const bool in_range = idx < a.P;
    const float* V = s_cam;
    const float* Mx = s_cam + 16;
    if (!a.early_params) stage_sh();

    // ---- per-Gaussian geometry ----
    unsigned tiles = 0;
    int rad = 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float px = 0, py = 0, pz = 0, depth = 0, ix = 0, iy = 0;
    float conA = 0, conB = 0, conC = 0, opac = 0;
    if (in_range) {
        px = a.means3D[3 * idx]; py = a.means3D[3 * idx + 1]; pz = a.means3D[3 * idx + 2];
        depth = xform_row(V, 2, px, py, pz);
        if (depth > 0.2f) {      // [upstream] in_frustum: p_view.z <= 0.2 culls
            float pvx = xform_row(V, 0, px, py, pz), pvy = xform_row(V, 1, px, py, pz);
            float phx = xform_row(Mx, 0, px, py, pz), phy = xform_row(Mx, 1, px, py, pz);
            float phw = xform_row(Mx, 3, px, py, pz);
            float pw = DIV(1.0f, ADD(phw, 0.0000001f));
            float ppx = MUL(phx, pw), ppy = MUL(phy, pw);
            float c3[6];
            if (a.cov3D_precomp) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = a.cov3D_precomp[6 * (size_t)idx + k];
            } else {
                float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
                cov3d_from_scale_rot(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2],
                                     a.scale_modifier, q.x, q.y, q.z, q.w, c3);
            }
            Cov2D cv;
            cov2d(pvx, pvy, depth, o.fx, o.fy, a.tanfovx, a.tanfovy, c3, V, cv);
            float det = FMA(cv.a, cv.c, -MUL(cv.b, cv.b));
            if (det != 0.0f) {
                float det_inv = DIV(1.0f, det);
                conA = MUL(cv.c, det_inv); conB = MUL(-cv.b, det_inv); conC = MUL(cv.a, det_inv);
                float mid = MUL(0.5f, ADD(cv.a, cv.c));
                float sq = SQRT(fmaxf(0.1f, FMA(mid, mid, -det)));
                float l1 = ADD(mid, sq), l2 = SUB(mid, sq);
                rad = __float2int_rz(ceilf(MUL(3.0f, SQRT(fmaxf(l1, l2)))));
                ix = MUL(FMA(ADD(ppx, 1.0f), (float)a.W, -1.0f), 0.5f);   // ndc2Pix
                iy = MUL(FMA(ADD(ppy, 1.0f), (float)a.H, -1.0f), 0.5f);
                get_rect(ix, iy, rad, o.gx, o.gy, x0, y0, x1, y1);
                tiles = (unsigned)((x1 - x0) * (y1 - y0));
                if (tiles == 0) rad = 0;
                opac = a.opacities[idx];
            }
        }
    }

    // ---- the Gaussian's sort item: depth bits (all its pairs share them), id, tile rectangle.
    //      A Gaussian without tiles sorts with key 0 and emits nothing.  The bits that differ
    //      among the VISIBLE keys are tracked so that a radix pass whose digit is the same for
    //      all of them (sign/exponent bytes of an avatar a few metres deep) can be skipped. ----
    const unsigned dkey = tiles ? __float_as_uint(depth) : 0u;
    if (in_range) {
        o.nkeys[idx] = dkey;
        o.nvals[idx] = (unsigned)idx;
        o.rects[idx] = make_uint2((unsigned)x0 | ((unsigned)y0 << 16),
                                  (unsigned)(x1 - x0) | ((unsigned)(y1 - y0) << 16));
        o.radii[idx] = rad;
#pragma unroll
        for (int p = 0; p < DEPTH_PASSES; p++)
            atomicAdd(&s_hist[p * RADIX + ((dkey >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    {
        const unsigned w_or = __reduce_or_sync(0xffffffffu, tiles ? dkey : 0u);
        const unsigned w_nor = __reduce_or_sync(0xffffffffu, tiles ? ~dkey : 0u);
        if ((tid & 31) == 0) {
            if (w_or) atomicOr(&s_var[0], w_or);
            if (w_nor) atomicOr(&s_var[1], w_nor);
        }
    }
    if constexpr (HAS_SH) cp_async_wait_all();
    __syncthreads();

    // ---- colour: SH -> RGB (or precomputed) and the blend record ----
    if (tiles > 0) {
        float rgb[3];
        unsigned flags = 0;
        if constexpr (HAS_SH) {
            float dx = SUB(px, V[32]), dy = SUB(py, V[33]), dz = SUB(pz, V[34]);
            float len = SQRT(FMA(dz, dz, FMA(dy, dy, MUL(dx, dx))));
            float inv = DIV(1.0f, len);
            float b[NB];
            sh_basis<D>(MUL(dx, inv), MUL(dy, inv), MUL(dz, inv), b);
            float sh[NVEC * 4];
            const float4* row = s_sh + tid * S4;
#pragma unroll
            for (int j = 0; j < NVEC; j++) {
                float4 v = row[j];
                sh[4 * j] = v.x; sh[4 * j + 1] = v.y; sh[4 * j + 2] = v.z; sh[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float acc = MUL(b[0], sh[ch]);
#pragma unroll
                for (int k = 1; k < NB; k++) acc = FMA(b[k], sh[3 * k + ch], acc);
                acc = ADD(acc, 0.5f);
                if (acc < 0.0f) flags |= 1u << ch;
                rgb[ch] = fmaxf(acc, 0.0f);
            }
        } else {
            rgb[0] = a.colors_precomp[3 * idx]; rgb[1] = a.colors_precomp[3 * idx + 1];
            rgb[2] = a.colors_precomp[3 * idx + 2];
        }
        // pmin: any power below it gives alpha = opacity*exp(power) < 1/255 with a 1 % margin,
        // so the blend kernels may skip the pair without evaluating exp (pure optimisation).
        float pmin = opac > 0.0f ? -(__logf(255.0f * opac) + 0.01f) : __int_as_float(0x7f800000);
        // footprint {alpha >= 1/255} = {d^T conic d <= t}, t = -2 pmin (with a margin); the blend
        // kernels compare it with the minimum of the quadratic form over a pixel block to skip
        // blocks the Gaussian cannot reach (exact ellipse/rectangle test).
        const float tq = -2.0f * pmin;
        const float t_m = (tq == tq) ? tq * 1.002f + 0.02f : __int_as_float(0x7f800000);   // NaN: no culling
        const float nb_c = -conB / conC, nb_a = -conB / conA;
        float4* rec = reinterpret_cast<float4*>(o.rec) + (size_t)idx * 4;
        rec[0] = make_float4(ix, iy, MUL(-0.5f, conA), -conB);
        rec[1] = make_float4(MUL(-0.5f, conC), opac, pmin, rgb[0]);
        rec[2] = make_float4(rgb[1], rgb[2], depth, 0.0f);
        rec[3] = make_float4(t_m, nb_c, nb_a, __uint_as_float(flags));
    }
    for (int i = tid; i < DEPTH_PASSES * RADIX; i += GEO_THREADS) {
        unsigned c = s_hist[i];
        if (c) atomicAdd(&o.hist[i], c);
    }
    if (tid < 2 && s_var[tid]) atomicOr(&o.varbits[tid], s_var[tid]);
}

// ------------------------------------------------------------------------------------------
// Pair emission in depth order: [upstream] InclusiveSum + duplicateWithKeys.  Thread i of
// chunk c takes the Gaussian at position c*256+i of the depth-sorted list.
// ------------------------------------------------------------------------------------------
struct EmitArgs {
    int P;
    const unsigned* nkeys[2];   // depth-sorted keys are in buffer depth_sort_parity(varbits)
    const unsigned* nvals[2];
    const unsigned* varbits;
    const uint2* rects;
    int* counters;
    unsigned* hist;             // [pass][256]; this kernel fills passes >= DEPTH_PASSES
    unsigned long long* scan_status;
    unsigned long long* keys;
    unsigned* vals;
    long long L_cap;
    int passes;
    int gx;
    int* host_counters;         // mapped host memory for {num_rendered, overflow}, or null
};

__global__ void __launch_bounds__(GEO_THREADS) emit_pairs_kernel(EmitArgs a) {
    __shared__ unsigned s_incl[GEO_THREADS];         // block-local inclusive tile offsets
    __shared__ int4 s_rect[GEO_THREADS];             // x0, y0, width, depth bits
    __shared__ unsigned s_gid[GEO_THREADS];
    __shared__ unsigned s_warp[GEO_THREADS / 32];
    __shared__ unsigned s_hist[(MAX_PASSES - DEPTH_PASSES) * RADIX];
    __shared__ int s_ticket;
    __shared__ unsigned long long s_wsum[GEO_THREADS / 32];
    __shared__ int s_wincl[GEO_THREADS / 32];
    __shared__ unsigned long long s_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_passes = a.passes - DEPTH_PASSES;
    pdl_sync();
    if (tid == 0) s_ticket = atomicAdd(&a.counters[CNT_SCAN_TICKET], 1);
    for (int i = tid; i < tile_passes * RADIX; i += GEO_THREADS) s_hist[i] = 0;
    __syncthreads();
    const int chunk = s_ticket;
    const int base = chunk * GEO_THREADS;
    const int pos_sorted = base + tid;
    const int par = depth_sort_parity(a.varbits, DEPTH_PASSES);
    unsigned tiles = 0;
    if (pos_sorted < a.P) {
        const unsigned dkey = (par ? a.nkeys[1] : a.nkeys[0])[pos_sorted];
        const unsigned gid = (par ? a.nvals[1] : a.nvals[0])[pos_sorted];
        const uint2 r = __ldg(a.rects + gid);
        const int w = (int)(r.y & 0xffffu), h = (int)(r.y >> 16);
        tiles = (unsigned)(w * h);
        s_rect[tid] = make_int4((int)(r.x & 0xffffu), (int)(r.x >> 16), w, (int)dkey);
        s_gid[tid] = gid;
    }
    unsigned incl = tiles;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < GEO_THREADS / 32; w++) {
        unsigned v = s_warp[w];
        if (w < warp) warp_off += v;
        block_total += v;
    }
    incl += warp_off;
    s_incl[tid] = incl;

#if SGS_EMIT_WIDE
    // ---- chained scan across CTAs: decoupled look-back by the WHOLE CTA, 256 predecessors
    // per round trip.  All emission CTAs are normally resident and publish their aggregates
    // at the same time, so a walk of 32 predecessors per round trip (one warp) would make
    // chunk k wait ~k/64 dependent L2 round trips; here it is ~k/512 + 1.
    if (tid == 0)
        st_relaxed_u64(&a.scan_status[chunk], (chunk == 0 ? FLAG_INCL : FLAG_AGG) | block_total);
    unsigned long long excl = 0;                     // CTA-uniform
    if (chunk > 0) {
        for (int look = chunk - 1;; look -= GEO_THREADS) {
            const int j = look - tid;
            unsigned long long s = j >= 0 ? ld_relaxed_u64(&a.scan_status[j]) : FLAG_INCL;
            while ((s & FLAG_MASK) == 0) s = ld_relaxed_u64(&a.scan_status[j]);
            const unsigned inc_mask = __ballot_sync(0xffffffffu, (s & FLAG_MASK) == FLAG_INCL);
            const int first = inc_mask ? (__ffs(inc_mask) - 1) : 31;
            unsigned long long v = lane <= first ? (s & ~FLAG_MASK) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) {
                s_wsum[warp] = v;
                s_wincl[warp] = inc_mask != 0;
            }
            __syncthreads();
            bool done = false;                       // warp 0 holds the nearest predecessors
#pragma unroll
            for (int w = 0; w < GEO_THREADS / 32; w++) {
                if (!done) {
                    excl += s_wsum[w];
                    done = s_wincl[w] != 0;
                }
            }
            __syncthreads();                         // before the next round overwrites s_wsum
            if (done) break;
        }
        if (tid == 0) st_relaxed_u64(&a.scan_status[chunk], FLAG_INCL | (excl + block_total));
    }
    if (tid == 0) {
        unsigned long long total = excl + block_total;
        if (chunk == (int)gridDim.x - 1) {
            a.counters[CNT_NUM_RENDERED] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
            if (a.host_counters) {     // the grand total decides the overflow: the last chunk knows both
                a.host_counters[0] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
                a.host_counters[1] = total > (unsigned long long)a.L_cap ? 1 : 0;
                __threadfence_system();
            }
        }
        if (total > (unsigned long long)a.L_cap) a.counters[CNT_OVERFLOW] = 1;
    }
    __syncthreads();          // s_incl, s_rect

    const unsigned long long prefix = excl;
#else
    // ---- chained scan across CTAs (decoupled look-back, one warp) ----
    if (warp == 0) {
        unsigned long long excl = 0;
        if (lane == 0)
            st_relaxed_u64(&a.scan_status[chunk], (chunk == 0 ? FLAG_INCL : FLAG_AGG) | block_total);
        if (chunk > 0) {
            int look = chunk - 1;
            while (true) {
                int j = look - lane;
                unsigned long long s = j >= 0 ? ld_relaxed_u64(&a.scan_status[j]) : FLAG_INCL;
                while (__any_sync(0xffffffffu, (s & FLAG_MASK) == 0)) {
                    if ((s & FLAG_MASK) == 0) s = ld_relaxed_u64(&a.scan_status[j]);
                }
                unsigned inc_mask = __ballot_sync(0xffffffffu, (s & FLAG_MASK) == FLAG_INCL);
                unsigned long long val = s & ~FLAG_MASK;
                int first = inc_mask ? (__ffs(inc_mask) - 1) : 31;
                unsigned long long v = lane <= first ? val : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(&a.scan_status[chunk], FLAG_INCL | (excl + block_total));
        }
        if (lane == 0) {
            s_prefix = excl;
            unsigned long long total = excl + block_total;
            if (chunk == (int)gridDim.x - 1) {
                a.counters[CNT_NUM_RENDERED] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
                if (a.host_counters) {     // the grand total decides the overflow: the last chunk knows both
                    a.host_counters[0] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
                    a.host_counters[1] = total > (unsigned long long)a.L_cap ? 1 : 0;
                    __threadfence_system();
                }
            }
            if (total > (unsigned long long)a.L_cap) a.counters[CNT_OVERFLOW] = 1;
        }
    }
    __syncthreads();          // s_prefix, s_incl, s_rect

    const unsigned long long prefix = s_prefix;
#endif
    // ---- emit (tile|depth) keys and Gaussian ids ----
    for (unsigned e = tid; e < block_total; e += GEO_THREADS) {
        int lo = 0, hi = GEO_THREADS - 1;          // first g with s_incl[g] > e
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_incl[mid] > e) hi = mid; else lo = mid + 1;
        }
        const int g = lo;
        const int4 r = s_rect[g];
        const unsigned first = g ? s_incl[g - 1] : 0u;
        const unsigned k = e - first;
        const unsigned ty = k / (unsigned)r.z, tx = k - ty * (unsigned)r.z;
        const unsigned tile_id = (unsigned)(r.y + (int)ty) * (unsigned)a.gx + (unsigned)(r.x + (int)tx);
        const unsigned long long key = ((unsigned long long)tile_id << 32) | (unsigned)r.w;
        const unsigned long long pos = prefix + e;
        if (pos < (unsigned long long)a.L_cap) {
            a.keys[pos] = key;
            a.vals[pos] = s_gid[g];
#if SGS_EMIT_MATCH
            // neighbouring pairs are neighbouring tiles: the upper digits are shared by most
            // of the warp, so equal digits are counted once (one shared atomic per group)
            const unsigned act = __activemask();
            for (int p = 0; p < tile_passes; p++) {
                const unsigned d = (tile_id >> (p * RADIX_BITS)) & (RADIX - 1);
                const unsigned peers = __match_any_sync(act, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[p * RADIX + d], (unsigned)__popc(peers));
            }
#else
            for (int p = 0; p < tile_passes; p++)
                atomicAdd(&s_hist[p * RADIX + ((tile_id >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
#endif
        }
    }
    __syncthreads();
    for (int i = tid; i < tile_passes * RADIX; i += GEO_THREADS) {
        unsigned c = s_hist[i];
        if (c) atomicAdd(&a.hist[DEPTH_PASSES * RADIX + i], c);
    }
}

int launch_emit_pairs(int P, const RasterLayout& lay, long long L_cap, char* bin, int* host_counters,
                      cudaStream_t stream) {
    if (P <= 0) return 0;
    EmitArgs a;
    a.P = P;
    a.nkeys[0] = reinterpret_cast<const unsigned*>(bin + lay.nkeys0_off);
    a.nkeys[1] = reinterpret_cast<const unsigned*>(bin + lay.nkeys1_off);
    a.nvals[0] = reinterpret_cast<const unsigned*>(bin + lay.nvals0_off);
    a.nvals[1] = reinterpret_cast<const unsigned*>(bin + lay.nvals1_off);
    a.counters = reinterpret_cast<int*>(bin + lay.cnt_off);
    a.varbits = reinterpret_cast<const unsigned*>(a.counters + CNT_VARBITS);
    a.rects = reinterpret_cast<const uint2*>(bin + lay.rects_off);
    a.hist = reinterpret_cast<unsigned*>(bin + lay.hist_off);
    a.scan_status = reinterpret_cast<unsigned long long*>(bin + lay.scan_off);
    a.keys = reinterpret_cast<unsigned long long*>(bin + lay.keys0_off);
    a.vals = reinterpret_cast<unsigned*>(bin + lay.vals0_off);
    a.L_cap = L_cap;
    a.passes = lay.passes;
    a.gx = lay.gx;
    a.host_counters = host_counters;
    launch_pdl(emit_pairs_kernel, lay.scan_blocks, GEO_THREADS, 0, stream, a);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// The pair list by count / scan / scatter.
//
// With the Gaussians in depth order (launch_depth_sort), the sorted (tile|depth, id) list is a
// STABLE counting sort of the pairs by tile id: tile t's segment holds, in depth order, the
// Gaussians whose rectangle covers t.  Three kernels replace emit_pairs + the tile-id radix
// passes and write the sorted list directly (bit-identical: a stable sort has one answer):
//   bin_count  : chunk c = BIN_GAUSS consecutive depth-ordered Gaussians -> counts[c][t]
//   bin_scan   : per tile, exclusive prefix of counts over the chunks -> base[c][t], total[t]
//   bin_scatter: tile starts (scan of total[], redone per CTA: a few thousand values), then the
//                chunk is enumerated again and every pair goes to start[t] + base[c][t] + rank.
// Enumeration (shared by count and scatter): a warp owns 128 consecutive Gaussians and walks
// their pairs in Gaussian-major order, 32 per step; the running per-(warp, tile) counters live
// in shared memory (u16), ranks inside a step come from match_any, so the order is stable by
// construction and no atomics are needed.  (Taking one Gaussian per step with lanes = its
// tiles was measured first: 33 us per pass -- 128 dependent steps per warp.)
// (A decoupled look-back does not carry over from the 256-bin radix passes: with one bin per
// tile the walk over predecessors would be 16x the traffic -- hence the separate scan.)
// ------------------------------------------------------------------------------------------
struct BinArgs {
    int P, tiles, gx, ctas;
    const unsigned* nkeys[2];   // depth-sorted keys are in buffer depth_sort_parity(varbits)
    const unsigned* nvals[2];
    const unsigned* varbits;
    const uint2* rects;
    unsigned short* counts;     // [ctas][tiles]
    unsigned* base;             // [ctas][tiles]
    unsigned* total;            // [tiles]
    int* counters;
    int* host_counters;         // mapped host memory for {num_rendered, overflow}, or null
    unsigned long long* keys;   // sorted list (output)
    unsigned* vals;
    long long L_cap;
};

// Per-warp staging of the enumeration: the warp's 128 Gaussians (4 consecutive ones per lane)
// with the inclusive prefix of their tile counts.
constexpr int BIN_WARP_GAUSS = BIN_GAUSS / (BIN_THREADS / 32);        // 128
struct BinWarpStage {
    unsigned pref[BIN_WARP_GAUSS];      // inclusive prefix of tiles touched
    uint4 info[BIN_WARP_GAUSS];         // x0 | y0 << 16, rect width, Gaussian id, depth bits
};

// PLACE = false: cnt[w][t] += 1 per pair.  PLACE = true: pair -> s_base[t] + cnt[w][t]++.
// The warp's pairs are taken 32 at a time in Gaussian-major order, one per lane (binary search
// of the staged prefix, like emit_pairs_kernel).  Pairs of one window that fall on the same
// tile come from different Gaussians and sit in lane order = depth order: match_any gives each
// its rank among them, the lowest lane advances the running per-(warp, tile) counter once.
template <bool PLACE>
__device__ __forceinline__ void bin_enumerate(const BinArgs& a, int chunk, unsigned short* s_cnt,
                                              const unsigned* s_base, BinWarpStage* s_stage) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int par = depth_sort_parity(a.varbits, DEPTH_PASSES);
    const unsigned* __restrict__ nk = par ? a.nkeys[1] : a.nkeys[0];
    const unsigned* __restrict__ nv = par ? a.nvals[1] : a.nvals[0];
    unsigned short* const cnt = s_cnt + (size_t)warp * a.tiles;
    BinWarpStage& st = s_stage[warp];
    const int g0 = chunk * BIN_GAUSS + warp * BIN_WARP_GAUSS + 4 * lane;
    // ---- stage: 4 consecutive Gaussians per lane ----
    unsigned n[4], tot = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int pos = g0 + i;
        unsigned gid = 0, dkey = 0;
        uint2 r = make_uint2(0u, 0u);
        if (pos < a.P) {
            gid = nv[pos];
            dkey = nk[pos];
            r = __ldg(a.rects + gid);
        }
        const unsigned rw = r.y & 0xffffu;
        n[i] = rw * (r.y >> 16);
        tot += n[i];
        st.info[4 * lane + i] = make_uint4(r.x, rw, gid, dkey);
    }
    unsigned incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    unsigned run = incl - tot;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        run += n[i];
        st.pref[4 * lane + i] = run;
    }
    const unsigned warp_total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    // ---- pairs, 32 per window ----
    for (unsigned e0 = 0; e0 < warp_total; e0 += 32) {
        const unsigned e = e0 + lane;
        const bool active = e < warp_total;
        unsigned t = 0xffff0000u | (unsigned)lane;      // inactive lanes: unique, matches nobody
        unsigned gid = 0, dkey = 0;
        if (active) {
            int lo = 0, hi = BIN_WARP_GAUSS - 1;        // first Gaussian whose inclusive prefix exceeds e
#pragma unroll
            for (int it = 0; it < 7; it++) {
                const int mid = (lo + hi) >> 1;
                if (st.pref[mid] > e) hi = mid; else lo = mid + 1;
            }
            const uint4 f = st.info[lo];
            const unsigned k = e - (lo ? st.pref[lo - 1] : 0u);
            const unsigned ty = k / f.y, tx = k - ty * f.y;
            t = ((f.x >> 16) + ty) * (unsigned)a.gx + (f.x & 0xffffu) + tx;
            gid = f.z; dkey = f.w;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, t);
        const int leader = __ffs(peers) - 1;
        unsigned c = 0;
        if (active && lane == leader) {
            c = cnt[t];
            cnt[t] = (unsigned short)(c + (unsigned)__popc(peers));
        }
        c = __shfl_sync(0xffffffffu, c, leader) + (unsigned)__popc(peers & lanemask_lt());
        if (PLACE && active) {
            const unsigned long long p = (unsigned long long)s_base[t] + c;
            if (p < (unsigned long long)a.L_cap) {
                a.keys[p] = ((unsigned long long)t << 32) | dkey;
                a.vals[p] = gid;
            }
        }
        __syncwarp();                                   // the next window may hit the same tiles
    }
    __syncwarp();                                       // the stage is rewritten by the next pass
}

__device__ __forceinline__ void bin_zero_counters(unsigned short* s_cnt, int tiles) {
    uint4* z = reinterpret_cast<uint4*>(s_cnt);         // 8 * tiles * 2 bytes, tiles even -> multiple of 16
    const int n16 = (BIN_THREADS / 32) * tiles * 2 / 16;
    for (int i = threadIdx.x; i < n16; i += BIN_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(BIN_THREADS) bin_count_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned char s_bin[];
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_bin);      // [8][tiles]
    BinWarpStage* s_stage = reinterpret_cast<BinWarpStage*>(s_bin + (size_t)(BIN_THREADS / 32) * a.tiles * 2);
    bin_zero_counters(s_cnt, a.tiles);
    pdl_sync();
    __syncthreads();
    const int chunk = blockIdx.x;
    bin_enumerate<false>(a, chunk, s_cnt, nullptr, s_stage);
    __syncthreads();
    // chunk totals per tile: two tiles per 32-bit word (counts <= 1024: no carry between halves)
    const unsigned* w32 = reinterpret_cast<const unsigned*>(s_cnt);
    unsigned* out = reinterpret_cast<unsigned*>(a.counts + (size_t)chunk * a.tiles);
    const int half = a.tiles / 2;
    for (int i = threadIdx.x; i < half; i += BIN_THREADS) {
        unsigned sum = 0;
#pragma unroll
        for (int w = 0; w < BIN_THREADS / 32; w++) sum += w32[(size_t)w * half + i];
        out[i] = sum;
    }
}

// Prefix over the chunks, per tile.  A CTA takes a slab of 32 tiles; its 8 warps split the chunk
// rows into 8 contiguous groups: sum the group (independent loads, all in flight), exchange the
// eight partial sums through shared memory, then walk the group again writing the prefixes.
// (One thread per tile walking all rows is a chain of dependent L2 round trips: measured 85 us.)
__global__ void __launch_bounds__(256) bin_scan_kernel(BinArgs a) {
    __shared__ unsigned s_part[8][32];
    pdl_sync();
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const bool ok = t < a.tiles;
    const int rows = (a.ctas + 7) / 8;
    const int r0 = min(a.ctas, grp * rows), r1 = min(a.ctas, r0 + rows);
    const unsigned short* __restrict__ cn = a.counts + (ok ? t : 0);
    unsigned* __restrict__ bs = a.base + (ok ? t : 0);
    const size_t stride = (size_t)a.tiles;
    unsigned sum = 0;
    if (ok) {
#pragma unroll 8
        for (int c = r0; c < r1; c++) sum += cn[(size_t)c * stride];
    }
    s_part[grp][lane] = sum;
    __syncthreads();
    unsigned run = 0;
#pragma unroll
    for (int g = 0; g < 8; g++)
        if (g < grp) run += s_part[g][lane];
    if (!ok) return;
    int c = r0;
    for (; c + 8 <= r1; c += 8) {               // loads first, then the dependent prefix and the stores
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = cn[(size_t)(c + k) * stride];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            bs[(size_t)(c + k) * stride] = run;
            run += v[k];
        }
    }
    for (; c < r1; c++) {
        const unsigned v = cn[(size_t)c * stride];
        bs[(size_t)c * stride] = run;
        run += v;
    }
    if (grp == 7) a.total[t] = run;             // the last group ends on the grand total
}

__global__ void __launch_bounds__(BIN_THREADS) bin_scatter_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned char s_bin[];
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_bin);                          // [8][tiles]
    unsigned* s_base = reinterpret_cast<unsigned*>(s_bin + (size_t)(BIN_THREADS / 32) * a.tiles * 2);   // [tiles]
    BinWarpStage* s_stage = reinterpret_cast<BinWarpStage*>(s_bin + (size_t)(BIN_THREADS / 32) * a.tiles * 2 + (size_t)a.tiles * 4);
    __shared__ unsigned s_wsum[BIN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bin_zero_counters(s_cnt, a.tiles);
    pdl_sync();
    const int chunk = blockIdx.x;
    // ---- tile starts: exclusive scan of total[] (every CTA redoes it: tiles <= 8192 values) ----
    const int per = (a.tiles + BIN_THREADS - 1) / BIN_THREADS;          // consecutive tiles per thread
    const int t0 = tid * per;
    unsigned mine = 0;
    for (int k = 0; k < per; k++)
        if (t0 + k < a.tiles) mine += a.total[t0 + k];
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    unsigned off = 0, grand = 0;
#pragma unroll
    for (int w = 0; w < BIN_THREADS / 32; w++) {
        const unsigned v = s_wsum[w];
        if (w < warp) off += v;
        grand += v;
    }
    unsigned start = off + incl - mine;
    const unsigned* brow = a.base + (size_t)chunk * a.tiles;
    for (int k = 0; k < per; k++) {
        if (t0 + k < a.tiles) {
            s_base[t0 + k] = start + brow[t0 + k];
            start += a.total[t0 + k];
        }
    }
    if (chunk == 0 && tid == 0) {
        const unsigned long long total = grand;
        a.counters[CNT_NUM_RENDERED] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
        if (total > (unsigned long long)a.L_cap) a.counters[CNT_OVERFLOW] = 1;
        if (a.host_counters) {
            a.host_counters[0] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
            a.host_counters[1] = total > (unsigned long long)a.L_cap ? 1 : 0;
            __threadfence_system();
        }
    }
    __syncthreads();
    // ---- pass A: per-warp counts; then exclusive over the warps (two tiles per word) ----
    bin_enumerate<false>(a, chunk, s_cnt, nullptr, s_stage);
    __syncthreads();
    unsigned* w32 = reinterpret_cast<unsigned*>(s_cnt);
    const int half = a.tiles / 2;
    for (int i = tid; i < half; i += BIN_THREADS) {
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < BIN_THREADS / 32; w++) {
            const unsigned v = w32[(size_t)w * half + i];
            w32[(size_t)w * half + i] = run;
            run += v;
        }
    }
    __syncthreads();
    // ---- pass B: place ----
    bin_enumerate<true>(a, chunk, s_cnt, s_base, s_stage);
}

bool bin_css_supported(const RasterLayout& lay) {
    return SGS_BIN_CSS && !SGS_MASKS_IN_SORT && lay.tiles <= BIN_MAX_TILES && (lay.tiles % 8) == 0;
}

int launch_bin_css(int P, const RasterLayout& lay, long long L_cap, char* bin, int* host_counters,
                   cudaStream_t stream, int debug) {
    if (P <= 0) return 0;
    BinArgs a;
    a.P = P; a.tiles = lay.tiles; a.gx = lay.gx; a.ctas = lay.bin_ctas;
    a.nkeys[0] = reinterpret_cast<const unsigned*>(bin + lay.nkeys0_off);
    a.nkeys[1] = reinterpret_cast<const unsigned*>(bin + lay.nkeys1_off);
    a.nvals[0] = reinterpret_cast<const unsigned*>(bin + lay.nvals0_off);
    a.nvals[1] = reinterpret_cast<const unsigned*>(bin + lay.nvals1_off);
    a.counters = reinterpret_cast<int*>(bin + lay.cnt_off);
    a.varbits = reinterpret_cast<const unsigned*>(a.counters + CNT_VARBITS);
    a.rects = reinterpret_cast<const uint2*>(bin + lay.rects_off);
    a.counts = reinterpret_cast<unsigned short*>(bin + lay.bcount_off);
    a.base = reinterpret_cast<unsigned*>(bin + lay.bbase_off);
    a.total = reinterpret_cast<unsigned*>(bin + lay.btotal_off);
    a.host_counters = host_counters;
    a.keys = reinterpret_cast<unsigned long long*>(bin + (lay.sorted_in_1() ? lay.keys1_off : lay.keys0_off));
    a.vals = reinterpret_cast<unsigned*>(bin + (lay.sorted_in_1() ? lay.vals1_off : lay.vals0_off));
    a.L_cap = L_cap;
    const size_t smem_stage = sizeof(BinWarpStage) * (BIN_THREADS / 32);
    const size_t smem_count = (size_t)(BIN_THREADS / 32) * lay.tiles * 2 + smem_stage;
    const size_t smem_scatter = (size_t)(BIN_THREADS / 32) * lay.tiles * 2 + (size_t)lay.tiles * 4 + smem_stage;
    SGS_CUDA_OK(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
    SGS_CUDA_OK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
    launch_pdl(bin_count_kernel, lay.bin_ctas, BIN_THREADS, smem_count, stream, a);
    SGS_LAUNCH_OK();
    SGS_STAGE_OK(debug, stream);
    launch_pdl(bin_scan_kernel, (lay.tiles + 31) / 32, 256, 0, stream, a);
    SGS_LAUNCH_OK();
    SGS_STAGE_OK(debug, stream);
    launch_pdl(bin_scatter_kernel, lay.bin_ctas, BIN_THREADS, smem_scatter, stream, a);
    SGS_LAUNCH_OK();
    SGS_STAGE_OK(debug, stream);
    return 0;
}

template <int D, bool HAS_SH>
static int launch_geo_t(const GeomArgs& a, const GeoOut& o, int blocks, bool vec16, cudaStream_t st) {
    size_t smem = HAS_SH ? (size_t)GEO_THREADS * sh_stride4(sh_nvec(D)) * 16 : 0;
    if (vec16) {
        auto k = geometry_kernel<D, HAS_SH, true>;
        if (smem > 48 * 1024) SGS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_pdl(k, blocks, GEO_THREADS, smem, st, a, o);
    } else {
        auto k = geometry_kernel<D, HAS_SH, false>;
        if (smem > 48 * 1024) SGS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_pdl(k, blocks, GEO_THREADS, smem, st, a, o);
    }
    SGS_LAUNCH_OK();
    return 0;
}

int launch_geometry(const GeomArgs& a, const RasterLayout& lay, long long L_cap, int* radii,
                    char* geom, char* bin, cudaStream_t stream, bool clear) {
    // one memset clears counters, histograms, scan + sort look-back status and the tile-length bucket counts
    if (clear) SGS_CUDA_OK(cudaMemsetAsync(bin, 0, lay.zero_bytes, stream));
    if (a.P <= 0) return 0;
    if (lay.gx > 0xffff || lay.gy > 0xffff) return SGS_ERR_CAPACITY;
    GeoOut o;
    o.radii = radii;
    o.rec = reinterpret_cast<float*>(geom + lay.rec_off);
    o.hist = reinterpret_cast<unsigned*>(bin + lay.hist_off);
    o.varbits = reinterpret_cast<unsigned*>(reinterpret_cast<int*>(bin + lay.cnt_off) + CNT_VARBITS);
    o.nkeys = reinterpret_cast<unsigned*>(bin + lay.nkeys0_off);
    o.nvals = reinterpret_cast<unsigned*>(bin + lay.nvals0_off);
    o.rects = reinterpret_cast<uint2*>(bin + lay.rects_off);
    o.gx = lay.gx; o.gy = lay.gy;
    o.fx = (float)a.W / (2.0f * a.tanfovx);
    o.fy = (float)a.H / (2.0f * a.tanfovy);
    const bool has_sh = a.colors_precomp == nullptr;
    if (!has_sh) return launch_geo_t<0, false>(a, o, lay.scan_blocks, false, stream);
    const bool vec16 = ((a.M * 3) % 4 == 0) && (((uintptr_t)a.shs & 15) == 0) &&
                       (a.M * 3 >= sh_nvec(a.D) * 4);
    switch (a.D) {
        case 0: return launch_geo_t<0, true>(a, o, lay.scan_blocks, vec16, stream);
        case 1: return launch_geo_t<1, true>(a, o, lay.scan_blocks, vec16, stream);
        case 2: return launch_geo_t<2, true>(a, o, lay.scan_blocks, vec16, stream);
        case 3: return launch_geo_t<3, true>(a, o, lay.scan_blocks, vec16, stream);
        default: return SGS_ERR_BAD_SH_DEGREE;
    }
}

// [upstream] rasterizer_impl.cu checkFrustum (markVisible)
__global__ void mark_visible_kernel(int P, const float* means3D, const float* view, unsigned char* present) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float z = xform_row(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    present[i] = z > 0.2f;
}

int launch_mark_visible(int P, const float* means3D, const float* view, unsigned char* present,
                        cudaStream_t stream) {
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, view, present);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs

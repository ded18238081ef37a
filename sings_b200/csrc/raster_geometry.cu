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
#include "lbs_math.cuh"

namespace sgs {

#ifndef SGS_GEO_THREADS
#define SGS_GEO_THREADS 256
#endif
constexpr int GEO_THREADS = SGS_GEO_THREADS;
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
    l.itemcnt_off = o;  o += align_up(32 * 4, 256);                       // work items of the backward blend per length bucket
    l.zero_bytes = o;
    size_t np = (size_t)(P > 0 ? P : 1);
    l.nkeys0_off = o;   o += align_up(np * 4, 256);
    l.nkeys1_off = o;   o += align_up(np * 4, 256);
    l.nvals0_off = o;   o += align_up(np * 4, 256);
    l.nvals1_off = o;   o += align_up(np * 4, 256);
    l.rects_off = o;    o += align_up(np * 8, 256);
    l.ranges_off = o;   o += align_up((size_t)l.tiles * 8, 256);
    l.bktlist_off = o;  o += align_up((size_t)32 * l.tiles * 4, 256);        // tiles by length bucket
    l.itemlist_off = o; o += align_up((size_t)32 * l.tiles * 8 * 16, 256);   // backward work items (tile * 8 + block, start, entries, -) by length bucket
    size_t cap = (size_t)(L_cap > 0 ? L_cap : 1);
    l.plane_entries = (cap + 31) / 32 * 32;
    l.blklist_off = o;  o += align_up(8 * l.plane_entries * 8, 256);
    l.keys0_off = o;    o += align_up(cap * 8, 256);
    l.keys1_off = o;    o += align_up(cap * 8, 256);
    l.vals0_off = o;    o += align_up(cap * 4, 256);
    l.vals1_off = o;    o += align_up(cap * 4, 256);
    l.bin_bytes = o;
    // image state
    size_t pix = (size_t)W * H;
    l.finalT_off = 0;
    l.ncontrib_off = align_up(pix * 4, 256);
    l.nblk_off = l.ncontrib_off + align_up(pix * 4, 256);
    l.img_bytes = l.nblk_off + align_up(pix * 4, 256);
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
// FUSE: the deform segment of SinGS.forward (sings_hybrid.py:398-419) runs in this kernel's
// prologue -- canonical attributes and packed skinning weights in, deformed mean / quaternion /
// scale straight into the projection math (and out to global memory once, for the backward and
// for inspection) -- instead of a separate LBS kernel writing them and this one reading them back.
template <int D, bool HAS_SH, bool VEC16, bool FUSE>
__global__ void __launch_bounds__(GEO_THREADS, FUSE ? (768 / GEO_THREADS) : 1)
geometry_kernel(GeomArgs a, GeoOut o, LbsFuse lf) {
    constexpr int NB = (D + 1) * (D + 1);
    constexpr int NVEC = HAS_SH ? sh_nvec(D) : 0;
    constexpr int S4 = HAS_SH ? sh_stride4(NVEC) : 0;
    extern __shared__ float4 s_sh[];                 // GEO_THREADS * S4 float4
    __shared__ float s_cam[36];
    __shared__ unsigned s_hist[DEPTH_PASSES * RADIX];
    __shared__ unsigned s_var[2];
    __shared__ float4 s_A[FUSE ? 64 * 3 : 1];        // rows 0..2 of the joint transforms

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
    const int idx = base + tid;
    const bool in_range = idx < a.P;
    if (a.early_params) stage_sh();
    // the canonical attributes are model parameters too: fetched ahead of the wait (the kernel in
    // front is pose -> A, which writes none of them)
    CanonG cg;
    if constexpr (FUSE) {
#pragma unroll
        for (int k = 0; k < 9; k++) cg.Rc[k] = (k % 4 == 0) ? 1.0f : 0.0f;
        if (in_range) load_canon(lf, idx, cg);
    }
    pdl_sync();
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    else if (tid < 37) s_var[tid - 35] = 0;
    for (int i = tid; i < DEPTH_PASSES * RADIX; i += GEO_THREADS) s_hist[i] = 0;
    if constexpr (FUSE) {
        for (int f = tid; f < lf.J * 3; f += GEO_THREADS)
            s_A[f] = reinterpret_cast<const float4*>(lf.A)[(f / 3) * 4 + f % 3];
    }
    __syncthreads();
    const float* V = s_cam;
    const float* Mx = s_cam + 16;
    if (!a.early_params) stage_sh();
    float m3[3] = {0, 0, 0}, sc3[3] = {0, 0, 0};
    float4 qrot = make_float4(1, 0, 0, 0);
    if constexpr (FUSE) {
        if (in_range) {
            const float sm = lf.smpl_scale ? __ldg(lf.smpl_scale) : 1.0f;
            float tr[3] = {0, 0, 0};
            if (lf.transl) { tr[0] = __ldg(lf.transl); tr[1] = __ldg(lf.transl + 1); tr[2] = __ldg(lf.transl + 2); }
            float q[4], T[12];
            deform_one(lf, s_A, cg, sm, tr, m3, q, sc3, T);
            qrot = make_float4(q[0], q[1], q[2], q[3]);
            lf.xyz_out[3 * (size_t)idx] = m3[0]; lf.xyz_out[3 * (size_t)idx + 1] = m3[1]; lf.xyz_out[3 * (size_t)idx + 2] = m3[2];
            reinterpret_cast<float4*>(lf.rotq_out)[idx] = qrot;
            lf.scales_out[3 * (size_t)idx] = sc3[0]; lf.scales_out[3 * (size_t)idx + 1] = sc3[1]; lf.scales_out[3 * (size_t)idx + 2] = sc3[2];
        }
    } else if (in_range) {
        m3[0] = a.means3D[3 * idx]; m3[1] = a.means3D[3 * idx + 1]; m3[2] = a.means3D[3 * idx + 2];
    }

    // ---- per-Gaussian geometry ----
    unsigned tiles = 0;
    int rad = 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float px = 0, py = 0, pz = 0, depth = 0, ix = 0, iy = 0;
    float conA = 0, conB = 0, conC = 0, opac = 0;
    if (in_range) {
        px = m3[0]; py = m3[1]; pz = m3[2];
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
                if constexpr (!FUSE) {
                    qrot = reinterpret_cast<const float4*>(a.rotations)[idx];
                    sc3[0] = a.scales[3 * idx]; sc3[1] = a.scales[3 * idx + 1]; sc3[2] = a.scales[3 * idx + 2];
                }
                cov3d_from_scale_rot(sc3[0], sc3[1], sc3[2], a.scale_modifier, qrot.x, qrot.y, qrot.z, qrot.w, c3);
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
    unsigned* vals;             // Gaussian id | reach mask << 24
    const float4* rec;          // geometry records (reach masks)
    long long L_cap;
    int passes;
    int gx;
    int* host_counters;         // mapped host memory for {num_rendered, overflow}, or null
};

__global__ void __launch_bounds__(GEO_THREADS) emit_pairs_kernel(EmitArgs a) {
    __shared__ unsigned s_incl[GEO_THREADS];         // block-local inclusive tile offsets
    __shared__ int4 s_rect[GEO_THREADS];             // x0, y0, width, depth bits
    __shared__ unsigned s_gid[GEO_THREADS];
    __shared__ float4 s_q0[GEO_THREADS], s_q3[GEO_THREADS];     // record parts the reach mask needs ...
    __shared__ float s_q1x[GEO_THREADS];                        // ... fetched once per Gaussian, not once per pair
    __shared__ unsigned s_warp[GEO_THREADS / 32];
    __shared__ unsigned s_hist[(MAX_PASSES - DEPTH_PASSES) * RADIX];
    __shared__ int s_ticket;
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
        if (tiles) {
            const float4* rp = a.rec + 4 * (size_t)gid;
            s_q0[tid] = __ldg(rp); s_q1x[tid] = __ldg(reinterpret_cast<const float*>(rp + 1)); s_q3[tid] = __ldg(rp + 3);
        }
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
                        }
            }
            if (total > (unsigned long long)a.L_cap) a.counters[CNT_OVERFLOW] = 1;
        }
    }
    __syncthreads();          // s_prefix, s_incl, s_rect

    const unsigned long long prefix = s_prefix;
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
            const float4 q1 = make_float4(s_q1x[g], 0.0f, 0.0f, 0.0f);
            const unsigned mask = reach_mask(s_q0[g], q1, s_q3[g], (float)((r.x + (int)tx) * TILE), (float)((r.y + (int)ty) * TILE));
            a.keys[pos] = key;
            a.vals[pos] = s_gid[g] | (mask << ID_BITS);
            for (int p = 0; p < tile_passes; p++)
                atomicAdd(&s_hist[p * RADIX + ((tile_id >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < tile_passes * RADIX; i += GEO_THREADS) {
        unsigned c = s_hist[i];
        if (c) atomicAdd(&a.hist[DEPTH_PASSES * RADIX + i], c);
    }
}

int launch_emit_pairs(int P, const RasterLayout& lay, long long L_cap, const char* geom, char* bin,
                      int* host_counters, cudaStream_t stream) {
    if (P <= 0) return 0;
    if (P > (int)ID_MASK) return SGS_ERR_CAPACITY;      // ids share their 32-bit word with the reach mask
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
    a.rec = reinterpret_cast<const float4*>(geom + lay.rec_off);
    a.L_cap = L_cap;
    a.passes = lay.passes;
    a.gx = lay.gx;
    a.host_counters = host_counters;
    launch_pdl(emit_pairs_kernel, lay.scan_blocks, GEO_THREADS, 0, stream, a);
    SGS_LAUNCH_OK();
    return 0;
}

template <int D, bool HAS_SH>
static int launch_geo_t(const GeomArgs& a, const GeoOut& o, int blocks, bool vec16, cudaStream_t st, const LbsFuse* lf) {
    size_t smem = HAS_SH ? (size_t)GEO_THREADS * sh_stride4(sh_nvec(D)) * 16 : 0;
    const LbsFuse none{};
    if (lf) {
        if constexpr (HAS_SH) {
            if (!vec16) return SGS_ERR_MISALIGNED;
            auto k = geometry_kernel<D, true, true, true>;
            SGS_CUDA_OK(set_max_smem(k, smem));
            SGS_CUDA_OK(launch_pdl(k, blocks, GEO_THREADS, smem, st, a, o, *lf));
            return 0;
        } else {
            return SGS_ERR_BAD_ARG;
        }
    }
    if (vec16) {
        auto k = geometry_kernel<D, HAS_SH, true, false>;
        SGS_CUDA_OK(set_max_smem(k, smem));
        SGS_CUDA_OK(launch_pdl(k, blocks, GEO_THREADS, smem, st, a, o, none));
    } else {
        auto k = geometry_kernel<D, HAS_SH, false, false>;
        SGS_CUDA_OK(set_max_smem(k, smem));
        SGS_CUDA_OK(launch_pdl(k, blocks, GEO_THREADS, smem, st, a, o, none));
    }
    return 0;
}

int launch_geometry(const GeomArgs& a, const RasterLayout& lay, long long L_cap, int* radii,
                    char* geom, char* bin, cudaStream_t stream, bool clear, const LbsFuse* lf) {
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
    if (!has_sh) return launch_geo_t<0, false>(a, o, lay.scan_blocks, false, stream, lf);
    const bool vec16 = ((a.M * 3) % 4 == 0) && (((uintptr_t)a.shs & 15) == 0) &&
                       (a.M * 3 >= sh_nvec(a.D) * 4);
    switch (a.D) {
        case 0: return launch_geo_t<0, true>(a, o, lay.scan_blocks, vec16, stream, lf);
        case 1: return launch_geo_t<1, true>(a, o, lay.scan_blocks, vec16, stream, lf);
        case 2: return launch_geo_t<2, true>(a, o, lay.scan_blocks, vec16, stream, lf);
        case 3: return launch_geo_t<3, true>(a, o, lay.scan_blocks, vec16, stream, lf);
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

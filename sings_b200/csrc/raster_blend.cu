// raster_blend.cu -- tile ranges, front-to-back alpha blending (forward) and its backward.
//
// Replaces, with identical results, the reference rasterizer's
//   [upstream] rasterizer_impl.cu identifyTileRanges, forward.cu renderCUDA,
//   backward.cu renderCUDA (K7, K8, K9 of section 2.4),
// reached from /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// One CTA per 16x16 tile, one thread per pixel.  B200-first structure (results unchanged):
//  * tiles are processed longest-list-first (an order array built by the last CTA of the
//    range kernel), so the few tiles with thousands of pairs do not form the tail;
//  * each warp owns an 8x4 pixel block; while staging a batch of 256 pairs every thread also
//    computes, for its pair, which of the 8 pixel blocks the Gaussian's alpha >= 1/255 footprint
//    can reach (ellipse bounding box AND bounding circle, conservative); each warp then
//    compacts the batch to the pairs that can touch ITS block and loops only over those;
//  * 48-byte geometry records are prefetched into registers one batch ahead and parked in a
//    double-buffered shared stage: one barrier per batch, broadcast LDS.128 in the loop;
//  * forward exits a tile as soon as every pixel has saturated (__syncthreads_count);
//  * the backward reduce-scatters the nine per-pixel partial gradients across the warp
//    (recursive halving: 14 shuffles instead of 45) and issues ONE reduction instruction
//    per warp and Gaussian (lanes 0..8 -> nine consecutive floats).
#include "common.cuh"
#include "kernels.h"

namespace sgs {

constexpr int ORDER_BUCKETS = 256;
constexpr int BWD_U = 2;      // pairs whose alpha is evaluated together in the backward blend
constexpr int FWD_U = 4;      // pairs evaluated together per pixel in the forward blend

// [upstream] identifyTileRanges: ranges[tile] = [start, end) in the sorted list (pre-zeroed).
// The last CTA to finish also builds `order`: tile ids sorted by descending list length
// (counting sort on length/16, ties in arbitrary order).
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const unsigned long long* __restrict__ keys, int* counters, long long n_cap,
                   uint2* ranges, unsigned* __restrict__ order, int tiles) {
    __shared__ unsigned s_cnt[ORDER_BUCKETS];
    __shared__ unsigned s_tmp[8];
    __shared__ int s_last;
    const long long n = min((long long)counters[CNT_NUM_RENDERED], n_cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        unsigned cur = (unsigned)(keys[i] >> 32);
        if (i == 0) ranges[cur].x = 0;
        else {
            unsigned prev = (unsigned)(keys[i - 1] >> 32);
            if (prev != cur) {
                ranges[prev].y = (unsigned)i;
                ranges[cur].x = (unsigned)i;
            }
        }
        if (i == n - 1) ranges[cur].y = (unsigned)n;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&counters[CNT_RANGES_DONE], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_cnt[tid] = 0;
    __syncthreads();
    // bucket of every tile: independent (unrolled) L2 loads, then the shared-memory counts
    constexpr int PER = 16;                       // tiles per thread per sweep (4096 tiles/sweep)
    const uint2* vr = ranges;
    for (int base = 0; base < tiles; base += 256 * PER) {
        unsigned bk[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int t = base + u * 256 + tid;
            uint2 rg = make_uint2(0, 0);
            if (t < tiles) rg = __ldcg(vr + t);
            bk[u] = ORDER_BUCKETS - 1 - min((rg.y - rg.x) >> 4, (unsigned)ORDER_BUCKETS - 1);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const bool ok = base + u * 256 + tid < tiles;
            const unsigned peers = __match_any_sync(0xffffffffu, ok ? bk[u] : 0xffffffffu);
            if (ok && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[bk[u]], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    // exclusive scan of the 256 bucket counts
    unsigned v = s_cnt[tid], incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    unsigned off = 0;
    for (int w = 0; w < warp; w++) off += s_tmp[w];
    __syncthreads();
    s_cnt[tid] = off + incl - v;
    __syncthreads();
    for (int base = 0; base < tiles; base += 256 * PER) {
        unsigned bk[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int t = base + u * 256 + tid;
            uint2 rg = make_uint2(0, 0);
            if (t < tiles) rg = __ldcg(vr + t);
            bk[u] = ORDER_BUCKETS - 1 - min((rg.y - rg.x) >> 4, (unsigned)ORDER_BUCKETS - 1);
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int t = base + u * 256 + tid;
            const bool ok = t < tiles;
            const unsigned peers = __match_any_sync(0xffffffffu, ok ? bk[u] : 0xffffffffu);
            const int leader = __ffs(peers) - 1;
            unsigned slot = 0;
            if (ok && lane == leader) slot = atomicAdd(&s_cnt[bk[u]], (unsigned)__popc(peers));
            slot = __shfl_sync(0xffffffffu, slot, leader);
            if (ok) order[slot + __popc(peers & lanemask_lt())] = (unsigned)t;
        }
    }
}

static inline const unsigned long long* sorted_keys(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned long long*>(bin + ((lay.passes & 1) ? lay.keys1_off : lay.keys0_off));
}
static inline const unsigned* sorted_vals(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned*>(bin + ((lay.passes & 1) ? lay.vals1_off : lay.vals0_off));
}

int launch_tile_ranges(const RasterLayout& lay, long long L_cap, char* bin, cudaStream_t stream) {
    int* counters = reinterpret_cast<int*>(bin + lay.cnt_off);
    long long blocks = (L_cap + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;          // grid-stride: few CTAs take the "last CTA" path
    tile_ranges_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
        sorted_keys(lay, bin), counters, L_cap, reinterpret_cast<uint2*>(bin + lay.ranges_off),
        reinterpret_cast<unsigned*>(bin + lay.order_off), lay.tiles);
    SGS_LAUNCH_OK();
    return 0;
}

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
    const int warp = tid >> 5, lane = tid & 31;
    lx = ((warp & 1) << 3) + (lane & 7);
    ly = ((warp >> 1) << 2) + (lane >> 3);
}

// Can the footprint {alpha >= 1/255} of a Gaussian reach the pixel block [bx0,bx1]x[by0,by1]?
// Exact test: the minimum over the rectangle of the quadratic form q(d) = a dx^2 + 2b dx dy +
// c dy^2 (d = pixel - centre; power = -q/2) is compared with t = -2 pmin (+ margin).  For a
// convex form the minimum lies at the centre (if inside) or on one of the two edges nearest
// to it; both edge minima are evaluated in closed form.  q3 = (t with margin, -b/c, -b/a,
// flags) comes from the geometry kernel.  A block that fails can only produce pairs the blend
// loop would reject, so skipping it changes nothing.
__device__ __forceinline__ bool reaches_block(const float4 q0, const float4 q1, const float4 q3,
                                              float bx0, float bx1, float by0, float by1) {
    const float a = -2.0f * q0.z, b = -q0.w, c = -2.0f * q1.x;
    const float X0 = bx0 - q0.x, X1 = bx1 - q0.x, Y0 = by0 - q0.y, Y1 = by1 - q0.y;
    const float cx = fminf(fmaxf(0.0f, X0), X1), cy = fminf(fmaxf(0.0f, Y0), Y1);
    const float dy1 = fminf(fmaxf(q3.y * cx, Y0), Y1);        // minimiser on the edge x = cx
    const float dx2 = fminf(fmaxf(q3.z * cy, X0), X1);        // minimiser on the edge y = cy
    const float qa = a * cx * cx + 2.0f * b * cx * dy1 + c * dy1 * dy1;
    const float qb = a * dx2 * dx2 + 2.0f * b * dx2 * cy + c * cy * cy;
    return fminf(qa, qb) <= q3.x;
}

struct Rec {
    float4 q0, q1, q2, q3;
};
__device__ __forceinline__ Rec load_rec(const float4* __restrict__ rec, unsigned id) {
    const float4* p = rec + 4 * (size_t)id;
    Rec r;
    r.q0 = __ldg(p); r.q1 = __ldg(p + 1); r.q2 = __ldg(p + 2); r.q3 = __ldg(p + 3);
    return r;
}

// ------------------------------------------------------------------------------------------
// forward.  Warp-autonomous: every warp streams the tile's depth-sorted list by itself, 32
// pairs at a time (one per lane, prefetched one chunk ahead), keeps the pairs that can reach
// its own 8x4 pixel block (ballot) and appends their records, compacted, to a warp-private
// ring in shared memory.  The ring is consumed FWD_U pairs at a time by a two-stage software
// pipeline: EVAL computes the alphas of the next FWD_U pairs (independent chains: falloff,
// exp, clamp) while COMPOSITE folds the previous FWD_U into the pixel state.  COMPOSITE is
// branch-free and its only serial dependency per pair is one multiply (T), one predicate
// (done) and the colour FMAs, so a lone warp on an SM -- the tail of the kernel is the tile
// with the longest list -- still retires a pair every few cycles.  No CTA barrier anywhere;
// a warp leaves as soon as its 32 pixels have saturated.
// ------------------------------------------------------------------------------------------
constexpr int RING_SLOTS = 64;     // >= 32 + FWD_U

struct FwdBatch {
    float al[FWD_U], om[FWD_U], r[FWD_U], g[FWD_U], b[FWD_U], z[FWD_U];
    unsigned pos[FWD_U];
};

__global__ void __launch_bounds__(TILE_PIX, 2)
blend_fwd_kernel(const uint2* __restrict__ ranges, const unsigned* __restrict__ order,
                 const unsigned* __restrict__ point_list, const float4* __restrict__ rec,
                 const float* __restrict__ bg, int W, int H, int gx_tiles,
                 float* __restrict__ out_color, float* __restrict__ final_T,
                 unsigned* __restrict__ n_contrib, float* __restrict__ out_alpha,
                 float* __restrict__ out_depth) {
    __shared__ float4 s_q0[TILE_PIX / 32][RING_SLOTS];     // x, y, -a/2, -b
    __shared__ float4 s_q1[TILE_PIX / 32][RING_SLOTS];     // -c/2, opacity, pmin, r
    __shared__ float4 s_q2[TILE_PIX / 32][RING_SLOTS];     // g, b, depth, list position + 1

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = order[blockIdx.x];
    const int tile_x = (int)(tile % (unsigned)gx_tiles), tile_y = (int)(tile / (unsigned)gx_tiles);
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = tile_x * TILE + lx, py = tile_y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float bx0 = (float)(tile_x * TILE + ((warp & 1) << 3)), bx1 = bx0 + 7.0f;
    const float by0 = (float)(tile_y * TILE + ((warp >> 1) << 2)), by1 = by0 + 3.0f;
    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    float4* const rq0 = s_q0[warp];
    float4* const rq1 = s_q1[warp];
    float4* const rq2 = s_q2[warp];

    bool done = !inside;
    float T = 1.0f, T_fin = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, Dacc = 0.0f;
    unsigned last = 0;
    unsigned head = 0, tail = 0;       // ring: consumed / produced pair counts (warp-uniform)

    FwdBatch cur;                      // the batch waiting to be composited (starts empty)
#pragma unroll
    for (int u = 0; u < FWD_U; u++) {
        cur.al[u] = 0.0f; cur.om[u] = 1.0f; cur.r[u] = cur.g[u] = cur.b[u] = cur.z[u] = 0.0f; cur.pos[u] = 0;
    }

    // alphas of ring entries [base, base + FWD_U); entries at or beyond `limit` are empty
    auto eval = [&](FwdBatch& e, unsigned base, unsigned limit) {
#pragma unroll
        for (int u = 0; u < FWD_U; u++) {
            const unsigned slot = (base + u) & (RING_SLOTS - 1);
            const float4 q0 = rq0[slot];
            const float4 q1 = rq1[slot];
            const float4 q2 = rq2[slot];
            const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
            const float uu = __fmul_rn(q0.z, dx), vv = __fmul_rn(q1.x, dy), ww = __fmul_rn(q0.w, dx);
            const float power = __fmaf_rn(ww, dy, __fmaf_rn(vv, dy, __fmul_rn(uu, dx)));
            const float al = fminf(0.99f, __fmul_rn(q1.y, expneg(fminf(power, 0.0f))));
            const bool ok = (base + u < limit) && !(power > 0.0f) && !(al < 1.0f / 255.0f);
            e.al[u] = ok ? al : 0.0f;
            e.om[u] = ok ? __fsub_rn(1.0f, al) : 1.0f;
            e.r[u] = q1.w; e.g[u] = q2.x; e.b[u] = q2.y; e.z[u] = q2.z;
            e.pos[u] = __float_as_uint(q2.w);
        }
    };
    // fold a batch into the pixel state, in order.  T keeps multiplying after saturation
    // (T_fin holds the reported value); an empty / rejected entry has al = 0, om = 1.
    auto composite = [&](const FwdBatch& e) {
#pragma unroll
        for (int u = 0; u < FWD_U; u++) {
            const float test_T = __fmul_rn(T, e.om[u]);
            const bool hit = e.al[u] > 0.0f;
            const bool kill = hit && !done && test_T < 0.0001f;
            T_fin = kill ? T : T_fin;
            done = done || kill;
            const float wgt = done ? 0.0f : __fmul_rn(e.al[u], T);
            C0 = __fmaf_rn(e.r[u], wgt, C0);
            C1 = __fmaf_rn(e.g[u], wgt, C1);
            C2 = __fmaf_rn(e.b[u], wgt, C2);
            Dacc = __fmaf_rn(e.z[u], wgt, Dacc);
            last = (hit && !done) ? e.pos[u] : last;
            T = test_T;
        }
    };

    // ring slots always hold finite records (an empty slot contributes colour * 0)
    rq0[lane] = rq0[lane + 32] = make_float4(0, 0, 0, 0);
    rq1[lane] = rq1[lane + 32] = make_float4(0, 0, 0, 0);
    rq2[lane] = rq2[lane + 32] = make_float4(0, 0, 0, 0);

    Rec p;
    p.q0 = p.q1 = p.q2 = p.q3 = make_float4(0, 0, 0, 0);
    if (lane < len) p = load_rec(rec, __ldg(point_list + range.x + lane));
    for (int pos = 0; pos < len; pos += 32) {
        if (__all_sync(0xffffffffu, done)) break;
        const bool rel = (pos + lane < len) && reaches_block(p.q0, p.q1, p.q3, bx0, bx1, by0, by1);
        const unsigned bits = __ballot_sync(0xffffffffu, rel);
        __syncwarp();                  // earlier ring reads are complete before slots are reused
        if (rel) {
            const unsigned slot = (tail + __popc(bits & lanemask_lt())) & (RING_SLOTS - 1);
            rq0[slot] = p.q0;
            rq1[slot] = p.q1;
            rq2[slot] = make_float4(p.q2.x, p.q2.y, p.q2.z, __uint_as_float((unsigned)(pos + lane + 1)));
        }
        tail += __popc(bits);
        __syncwarp();
        if (pos + 32 + lane < len) p = load_rec(rec, __ldg(point_list + range.x + pos + 32 + lane));
        while (tail - head >= FWD_U) {
            FwdBatch nxt;
            eval(nxt, head, tail);
            composite(cur);
            cur = nxt;
            head += FWD_U;
        }
    }
    {   // drain: the waiting batch, then the (partial) remainder of the ring
        FwdBatch nxt;
        eval(nxt, head, tail);
        composite(cur);
        composite(nxt);
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
        const float Tr = done ? T_fin : T;       // a saturated pixel reports the T it stopped at
        final_T[pix] = Tr;
        n_contrib[pix] = last;
        out_color[pix] = __fmaf_rn(Tr, bg[0], C0);
        out_color[plane + pix] = __fmaf_rn(Tr, bg[1], C1);
        out_color[2 * plane + pix] = __fmaf_rn(Tr, bg[2], C2);
        if (out_alpha) out_alpha[pix] = __fsub_rn(1.0f, Tr);
        if (out_depth) out_depth[pix] = Dacc;
    }
}

int launch_blend_fwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     char* img, const float* bg, float* out_color, float* out_alpha,
                     float* out_depth, cudaStream_t stream) {
    blend_fwd_kernel<<<lay.tiles, TILE_PIX, 0, stream>>>(
        reinterpret_cast<const uint2*>(bin + lay.ranges_off),
        reinterpret_cast<const unsigned*>(bin + lay.order_off), sorted_vals(lay, bin),
        reinterpret_cast<const float4*>(geom + lay.rec_off), bg, W, H, lay.gx, out_color,
        reinterpret_cast<float*>(img + lay.finalT_off),
        reinterpret_cast<unsigned*>(img + lay.ncontrib_off), out_alpha, out_depth);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward.  Same warp-autonomous streaming, back to front, starting at the largest
// contributor count among the warp's 32 pixels.  The nine per-pixel partial gradients of a
// Gaussian are reduce-scattered across the warp (recursive halving: 14 shuffles instead of
// 45) and leave as ONE reduction instruction (lanes 0..8 -> nine consecutive floats).
// ------------------------------------------------------------------------------------------
// Sum eight values across the warp by recursive halving; lane l returns the total of v[l & 7].
__device__ __forceinline__ float reduce_scatter8(float v0, float v1, float v2, float v3, float v4,
                                                 float v5, float v6, float v7, int lane) {
    const bool h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
    float k0 = h4 ? v4 : v0, k1 = h4 ? v5 : v1, k2 = h4 ? v6 : v2, k3 = h4 ? v7 : v3;
    k0 += __shfl_xor_sync(0xffffffffu, h4 ? v0 : v4, 4);
    k1 += __shfl_xor_sync(0xffffffffu, h4 ? v1 : v5, 4);
    k2 += __shfl_xor_sync(0xffffffffu, h4 ? v2 : v6, 4);
    k3 += __shfl_xor_sync(0xffffffffu, h4 ? v3 : v7, 4);
    float m0 = h2 ? k2 : k0, m1 = h2 ? k3 : k1;
    m0 += __shfl_xor_sync(0xffffffffu, h2 ? k0 : k2, 2);
    m1 += __shfl_xor_sync(0xffffffffu, h2 ? k1 : k3, 2);
    float r = h1 ? m1 : m0;
    r += __shfl_xor_sync(0xffffffffu, h1 ? m0 : m1, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

__global__ void __launch_bounds__(TILE_PIX)
blend_bwd_kernel(const uint2* __restrict__ ranges, const unsigned* __restrict__ order,
                 const unsigned* __restrict__ point_list, const float4* __restrict__ rec,
                 const float* __restrict__ bg, int W, int H, int gx_tiles,
                 const float* __restrict__ final_T, const unsigned* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    __shared__ float4 s_q0[TILE_PIX / 32][2][32];
    __shared__ float4 s_q1[TILE_PIX / 32][2][32];
    __shared__ float4 s_q2[TILE_PIX / 32][2][32];
    __shared__ unsigned s_id[TILE_PIX / 32][2][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = order[blockIdx.x];
    const int tile_x = (int)(tile % (unsigned)gx_tiles), tile_y = (int)(tile / (unsigned)gx_tiles);
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = tile_x * TILE + lx, py = tile_y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float bx0 = (float)(tile_x * TILE + ((warp & 1) << 3)), bx1 = bx0 + 7.0f;
    const float by0 = (float)(tile_y * TILE + ((warp >> 1) << 2)), by1 = by0 + 3.0f;
    const uint2 range = ranges[tile];
    if (range.y == range.x) return;
    const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;

    const unsigned last = inside ? n_contrib[pix] : 0u;
    // entries at list positions >= the warp's largest contributor count reach none of its pixels
    const int wlast = (int)__reduce_max_sync(0xffffffffu, last);
    if (wlast == 0) return;
    const float T_final = inside ? final_T[pix] : 0.0f;
    float T = T_final;
    float dp0 = 0.0f, dp1 = 0.0f, dp2 = 0.0f;
    if (inside) { dp0 = dL_dpix[pix]; dp1 = dL_dpix[plane + pix]; dp2 = dL_dpix[2 * plane + pix]; }
    const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;

    float acc_r0 = 0, acc_r1 = 0, acc_r2 = 0;       // accum_rec
    float last_alpha = 0, lc0 = 0, lc1 = 0, lc2 = 0;

    Rec p;
    unsigned pid = 0;
    p.q0 = p.q1 = p.q2 = p.q3 = make_float4(0, 0, 0, 0);
    if (lane < wlast) {
        pid = __ldg(point_list + range.x + (wlast - 1 - lane));
        p = load_rec(rec, pid);
    }
    int buf = 0;
    // chunk c covers list positions top-lane, top = wlast-1-32c: lane order = back-to-front order
    for (int top = wlast - 1; top >= 0; top -= 32, buf ^= 1) {
        const bool rel = (top - lane >= 0) && reaches_block(p.q0, p.q1, p.q3, bx0, bx1, by0, by1);
        unsigned bits = __ballot_sync(0xffffffffu, rel);
        s_q0[warp][buf][lane] = p.q0; s_q1[warp][buf][lane] = p.q1; s_q2[warp][buf][lane] = p.q2;
        s_id[warp][buf][lane] = pid;
        __syncwarp();
        if (top - 32 - lane >= 0) {
            pid = __ldg(point_list + range.x + (top - 32 - lane));
            p = load_rec(rec, pid);
        }
        while (bits) {
            // stage 1: falloff and alpha of BWD_U pairs, independent chains
            float Gs[BWD_U], als[BWD_U], dxs[BWD_U], dys[BWD_U];
            int jj[BWD_U];
            bool vs[BWD_U];
            bool any_ok = false;
#pragma unroll
            for (int u = 0; u < BWD_U; u++) {
                const bool in = bits != 0;
                const int j = in ? __ffs(bits) - 1 : 0;
                bits &= bits - 1;
                const unsigned pos = (unsigned)(top - j);            // 0-based list position
                const float4 q0 = s_q0[warp][buf][j];
                const float4 q1 = s_q1[warp][buf][j];
                const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
                const float uu = __fmul_rn(q0.z, dx), vv = __fmul_rn(q1.x, dy), ww = __fmul_rn(q0.w, dx);
                const float power = __fmaf_rn(ww, dy, __fmaf_rn(vv, dy, __fmul_rn(uu, dx)));
                vs[u] = in && pos < last && !(power > 0.0f) && !(power < q1.z);
                Gs[u] = fminf(power, 0.0f);
                als[u] = q1.y;
                dxs[u] = dx; dys[u] = dy; jj[u] = j;
                any_ok |= vs[u];
            }
            if (!__any_sync(0xffffffffu, any_ok)) continue;
#pragma unroll
            for (int u = 0; u < BWD_U; u++) {
                Gs[u] = expneg(Gs[u]);
                als[u] = fminf(0.99f, __fmul_rn(als[u], Gs[u]));
                vs[u] = vs[u] && !(als[u] < 1.0f / 255.0f);
            }
            // stage 2: order-dependent state update, gradients, warp reduction
#pragma unroll
            for (int u = 0; u < BWD_U; u++) {
                const bool valid = vs[u];
                if (!__any_sync(0xffffffffu, valid)) continue;
                const int j = jj[u];
                float g_mx = 0, g_my = 0, g_ca = 0, g_cb = 0, g_cc = 0, g_op = 0, g_r = 0, g_g = 0, g_b = 0;
                if (valid) {
                    const float4 q0 = s_q0[warp][buf][j];
                    const float4 q1 = s_q1[warp][buf][j];
                    const float4 q2 = s_q2[warp][buf][j];
                    const float G = Gs[u], alpha = als[u], dx = dxs[u], dy = dys[u];
                    const float inv1a = 1.0f / (1.0f - alpha);
                    T = T * inv1a;
                    const float dch = alpha * T;
                    const float c0 = q1.w, c1 = q2.x, c2 = q2.y;
                    acc_r0 = last_alpha * lc0 + (1.0f - last_alpha) * acc_r0;
                    acc_r1 = last_alpha * lc1 + (1.0f - last_alpha) * acc_r1;
                    acc_r2 = last_alpha * lc2 + (1.0f - last_alpha) * acc_r2;
                    lc0 = c0; lc1 = c1; lc2 = c2;
                    float dL_dalpha = (c0 - acc_r0) * dp0 + (c1 - acc_r1) * dp1 + (c2 - acc_r2) * dp2;
                    g_r = dch * dp0; g_g = dch * dp1; g_b = dch * dp2;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final * inv1a) * bg_dot;
                    const float dL_dG = q1.y * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    // conic entries: a = -2*q0.z, b = -q0.w, c = -2*q1.x
                    const float ca = -2.0f * q0.z, cb = -q0.w, cc = -2.0f * q1.x;
                    const float dG_ddelx = -gdx * ca - gdy * cb;
                    const float dG_ddely = -gdy * cc - gdx * cb;
                    g_mx = dL_dG * dG_ddelx * ddelx_dx;
                    g_my = dL_dG * dG_ddely * ddely_dy;
                    g_ca = -0.5f * gdx * dx * dL_dG;
                    g_cb = -0.5f * gdx * dy * dL_dG;
                    g_cc = -0.5f * gdy * dy * dL_dG;
                    g_op = G * dL_dalpha;
                }
                // accumulator slots 0..8: mean2D.x, .y, conic a, b, c, opacity, r, g, b
                const float r8 = reduce_scatter8(g_mx, g_my, g_ca, g_cb, g_cc, g_op, g_r, g_g, lane);
                const float rb = warp_sum(g_b);
                if (lane < 9) atomicAdd(acc + (size_t)s_id[warp][buf][j] * ACC_FLOATS + lane, lane < 8 ? r8 : rb);
            }
        }
    }
}

int launch_blend_bwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     const char* img, const float* bg, const float* dL_dpix, float* acc,
                     cudaStream_t stream) {
    blend_bwd_kernel<<<lay.tiles, TILE_PIX, 0, stream>>>(
        reinterpret_cast<const uint2*>(bin + lay.ranges_off),
        reinterpret_cast<const unsigned*>(bin + lay.order_off), sorted_vals(lay, bin),
        reinterpret_cast<const float4*>(geom + lay.rec_off), bg, W, H, lay.gx,
        reinterpret_cast<const float*>(img + lay.finalT_off),
        reinterpret_cast<const unsigned*>(img + lay.ncontrib_off), dL_dpix, acc);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs

// raster_blend.cu -- tile ranges, front-to-back alpha blending (forward) and its backward.
//
// Replaces, with identical results, the reference rasterizer's
//   [upstream] rasterizer_impl.cu identifyTileRanges, forward.cu renderCUDA,
//   backward.cu renderCUDA (K7, K8, K9 of section 2.4),
// reached from /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// A warp per 8x4 pixel block, a thread per pixel, warps autonomous (no CTA barrier anywhere).
// B200-first structure (results unchanged):
//  * tiles / pixel blocks are processed longest-first (work items filed in buckets by length);
//  * a list entry carries, next to the Gaussian id, the mask of the tile's eight pixel blocks the
//    Gaussian's alpha >= 1/255 footprint can reach; the forward keeps only the entries of its
//    own block, and writes them out as the block's own list for the backward;
//  * 48-byte geometry records are gathered one chunk ahead into a warp-private ring in shared
//    memory whose geometric part is pair-interleaved: EVAL works on two entries per instruction
//    (packed binary32 pairs), COMPOSITE / SEQ are branch-free;
//  * the backward parks the two order-dependent numbers per (pixel, pair) in a [pair][pixel]
//    tile and reduces over pixels with lane = pair: 2 vector + 1 scalar reduction per warp and
//    Gaussian.
#include "common.cuh"
#include "kernels.h"
#include <type_traits>

namespace sgs {

// A/B knobs (tools/sweep.sh builds variants with -DSGS_...): pipeline depth vs. resident CTAs
#ifndef SGS_BWD_U
#define SGS_BWD_U 4
#endif
#ifndef SGS_FWD_U
#define SGS_FWD_U 4
#endif
#ifndef SGS_FWD_MINB
#define SGS_FWD_MINB 2
#endif
#ifndef SGS_BWD_MINB
#define SGS_BWD_MINB 2
#endif
#ifndef SGS_FWD_WPC              // warps per CTA of the blend kernels: the 8 warps of a tile are autonomous,
#define SGS_FWD_WPC 4            // so a tile may be spread over 8 / WPC smaller CTAs (finer scheduling grain)
#endif
#ifndef SGS_BWD_WPC
#define SGS_BWD_WPC 4
#endif
constexpr int FWD_WPC = SGS_FWD_WPC, BWD_WPC = SGS_BWD_WPC;
constexpr int TILE_WARPS = TILE_PIX / 32;
static_assert(TILE_WARPS % FWD_WPC == 0 && TILE_WARPS % BWD_WPC == 0, "warps per CTA must divide 8");
constexpr int BWD_U = SGS_BWD_U;      // pairs per software-pipeline stage in the backward blend
constexpr int FWD_U = SGS_FWD_U;      // pairs evaluated together per pixel in the forward blend

// [upstream] identifyTileRanges: ranges[tile] = [start, end) in the sorted list, (0, 0) for a
// tile without pairs.  One WARP per tile: a 33-ary search (32 probes per step, ballot) for the
// lower bounds of tile and tile + 1 over the sorted keys -- 4 dependent L2 round trips at a
// million pairs instead of the 21 of a binary search; no pass over the list, no atomics on
// the ranges, no memset.  The CTA (8 tiles) then files its tiles in buckets by list length
// (two buckets per octave of the length, bucket 0 = empty) so the blend kernels can take tiles
// longest-first: the few tiles with thousands of pairs must not form the tail.
constexpr int LEN_BUCKETS = 32;
constexpr int RANGE_THREADS = 256;

// bucket of a list of c >= 1 entries: two per octave, monotonic in c
__device__ __forceinline__ unsigned length_bucket(unsigned c) {
    const unsigned l = 31u - (unsigned)__clz(c);
    return l == 0u ? 0u : min((unsigned)LEN_BUCKETS - 1u, 2u * l + ((c >> (l - 1u)) & 1u));
}

// first index in [0, n) whose key's tile id is >= target (warp-cooperative)
__device__ __forceinline__ unsigned warp_lower_bound(const unsigned long long* __restrict__ keys,
                                                     unsigned n, unsigned target, int lane) {
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned len = hi - lo;
        if (len <= 32) {
            const bool pred = (unsigned)lane < len && (unsigned)(__ldg(keys + lo + lane) >> 32) < target;
            return lo + __popc(__ballot_sync(0xffffffffu, pred));
        }
        auto probe = [&](unsigned k) { return lo + (unsigned)(((unsigned long long)(k + 1) * len) / 33u); };
        const bool pred = (unsigned)(__ldg(keys + probe(lane)) >> 32) < target;
        const unsigned cnt = __popc(__ballot_sync(0xffffffffu, pred));      // keys are sorted: a prefix of lanes
        const unsigned nlo = cnt ? probe(cnt - 1) + 1 : lo;
        const unsigned nhi = cnt < 32 ? probe(cnt) : hi;
        lo = nlo; hi = nhi;
    }
    return lo;
}

__device__ __forceinline__ void tile_ranges_block(int block, const unsigned long long* __restrict__ keys,
                                                  const int* __restrict__ counters, long long n_cap,
                                                  uint2* __restrict__ ranges, unsigned* __restrict__ bucket_count,
                                                  unsigned* __restrict__ bucket_list, int tiles) {
    __shared__ unsigned s_len[RANGE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = block * (RANGE_THREADS / 32) + warp;
    const unsigned n = (unsigned)min((long long)counters[CNT_NUM_RENDERED], n_cap);
    unsigned c = 0;
    if (t < tiles) {                                     // warp-uniform
        const unsigned lo = warp_lower_bound(keys, n, (unsigned)t, lane);
        const unsigned hi = warp_lower_bound(keys, n, (unsigned)t + 1u, lane);
        c = hi - lo;
        if (lane == 0) ranges[t] = c ? make_uint2(lo, hi) : make_uint2(0u, 0u);
    }
    if (lane == 0) s_len[warp] = c;
    __syncthreads();
    if (warp != 0) return;
    const int tt = block * (RANGE_THREADS / 32) + lane;
    const bool ok = lane < RANGE_THREADS / 32 && tt < tiles;
    // bucket 0 = empty tiles (last in the longest-first order), else two buckets per octave of the length
    const unsigned bk = ok ? (s_len[lane] ? max(1u, length_bucket(s_len[lane])) : 0u) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, bk);
    const int leader = __ffs(peers) - 1;
    unsigned slot = 0;
    if (ok && lane == leader) slot = atomicAdd(&bucket_count[bk], (unsigned)__popc(peers));
    slot = __shfl_sync(0xffffffffu, slot, leader);
    if (ok) bucket_list[(size_t)bk * tiles + slot + __popc(peers & lanemask_lt())] = (unsigned)tt;
}

// The i-th tile in longest-bucket-first order (whole warp calls it with the same i).  The bucket
// prefix sums depend only on the frame, so a persistent warp builds them once (BucketScan: lane l
// holds bucket LEN_BUCKETS-1-l, lane 0 = longest) and a rank costs one load afterwards.
struct BucketScan {
    unsigned cnt, incl;
};
__device__ __forceinline__ BucketScan bucket_scan(const unsigned* __restrict__ bucket_count) {
    const int lane = threadIdx.x & 31;
    BucketScan b;
    b.cnt = __ldg(bucket_count + (LEN_BUCKETS - 1 - lane));
    b.incl = b.cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, b.incl, d);
        if (lane >= d) b.incl += x;
    }
    return b;
}
__device__ __forceinline__ unsigned tile_of_rank(const BucketScan& bs, const unsigned* __restrict__ bucket_list,
                                                 int tiles, unsigned i) {
    const unsigned before = __ballot_sync(0xffffffffu, bs.incl <= i);       // buckets entirely before i
    const int b = __popc(before);                                           // lane holding the bucket of i
    const unsigned excl = __shfl_sync(0xffffffffu, bs.incl - bs.cnt, b & 31);
    return __ldg(bucket_list + (size_t)(LEN_BUCKETS - 1 - b) * tiles + (i - excl));
}

// Work distribution of the two blend kernels: a tile is spread over 8 / WPC CTAs of WPC
// autonomous warps (warp = 8x4 pixel block), CTAs in longest-list-first order.  (Persistent
// warps / CTAs pulling items from a ticket counter were measured 10 % slower -- a hardware-
// scheduled CTA thins out as its short warps retire, which speeds up its long ones;
// profiles/README.md, round 1.)
template <int WPC, typename Fn>
__device__ __forceinline__ void for_each_tile_item(const unsigned* __restrict__ bucket_count,
                                              const unsigned* __restrict__ bucket_list, int tiles, Fn fn) {
    constexpr int PARTS = TILE_WARPS / WPC;      // CTAs per tile
    const BucketScan bs = bucket_scan(bucket_count);
    fn(tile_of_rank(bs, bucket_list, tiles, blockIdx.x / PARTS),
       (int)(blockIdx.x % PARTS) * WPC + (int)(threadIdx.x >> 5));
}

__device__ __forceinline__ uint4 item_of_rank(const BucketScan& bs, const uint4* __restrict__ bucket_list,
                                              size_t items_cap, unsigned i) {
    const unsigned before = __ballot_sync(0xffffffffu, bs.incl <= i);       // buckets entirely before i
    const int b = __popc(before);                                           // lane holding the bucket of i
    const unsigned excl = __shfl_sync(0xffffffffu, bs.incl - bs.cnt, b & 31);
    return __ldg(bucket_list + (size_t)(LEN_BUCKETS - 1 - b) * items_cap + (i - excl));
}

// Work items of the backward blend: one (tile, pixel block) whose pixels have at least one
// contributor, filed by the forward warp that rendered it in a bucket by the EXACT number of list
// entries the backward will walk (two buckets per octave), a warp per item, the four warps of a
// CTA taking four consecutive items of the longest-first order (similar lengths: the CTA's warps
// retire together).  The grid holds a quarter of the possible items (GRID_FOLD) -- most pixel
// blocks of an avatar frame are empty -- and warps stride over the items that are really there.
#ifndef SGS_GRID_FOLD
#define SGS_GRID_FOLD 4
#endif
constexpr int GRID_FOLD = SGS_GRID_FOLD;
constexpr int BLOCKS_PER_TILE = TILE_PIX / 32;            // 8 pixel blocks of 8x4
template <typename Fn>
__device__ __forceinline__ void for_each_item(const unsigned* __restrict__ bucket_count,
                                              const uint4* __restrict__ bucket_list, int tiles, Fn fn) {
    const BucketScan bs = bucket_scan(bucket_count);
    const unsigned n_items = __shfl_sync(0xffffffffu, bs.incl, 31);
    const unsigned wpc = blockDim.x >> 5;
    for (unsigned it = blockIdx.x * wpc + (threadIdx.x >> 5); it < n_items; it += gridDim.x * wpc)
        fn(item_of_rank(bs, bucket_list, (size_t)tiles * 8u, it));
}
template <int WPC>
static inline unsigned blend_grid(int tiles) {
    const long long items = (long long)tiles * BLOCKS_PER_TILE;
    return (unsigned)((items + (long long)WPC * GRID_FOLD - 1) / ((long long)WPC * GRID_FOLD));
}

static inline const unsigned long long* sorted_keys(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned long long*>(bin + (lay.sorted_in_1() ? lay.keys1_off : lay.keys0_off));
}
static inline const unsigned* sorted_vals(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned*>(bin + (lay.sorted_in_1() ? lay.vals1_off : lay.vals0_off));
}

__device__ __forceinline__ void pixel_of_thread(int warp, int lane, int& lx, int& ly) {
    lx = ((warp & 1) << 3) + (lane & 7);
    ly = ((warp >> 1) << 2) + (lane >> 3);
}

struct Rec {
    float4 q0, q1, q2;
};
__device__ __forceinline__ Rec load_rec(const float4* __restrict__ rec, unsigned id) {
    const float4* p = rec + 4 * (size_t)id;
    Rec r;
    r.q0 = __ldg(p); r.q1 = __ldg(p + 1); r.q2 = __ldg(p + 2);
    return r;
}

// [upstream] identifyTileRanges (+ the length buckets): 8 tiles per CTA, a warp each.
__global__ void __launch_bounds__(RANGE_THREADS)
tile_ranges_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ counters, long long n_cap,
                   int tiles, uint2* __restrict__ ranges, unsigned* __restrict__ bucket_count,
                   unsigned* __restrict__ bucket_list) {
    pdl_sync();
    tile_ranges_block((int)blockIdx.x, keys, counters, n_cap, ranges, bucket_count, bucket_list, tiles);
}

int launch_tile_ranges(const RasterLayout& lay, long long L_cap, char* bin, cudaStream_t stream) {
    const int range_blocks = (lay.tiles + RANGE_THREADS / 32 - 1) / (RANGE_THREADS / 32);
    SGS_CUDA_OK(launch_pdl(tile_ranges_kernel, (unsigned)range_blocks, RANGE_THREADS, 0, stream,
        sorted_keys(lay, bin), reinterpret_cast<const int*>(bin + lay.cnt_off), L_cap, lay.tiles,
        reinterpret_cast<uint2*>(bin + lay.ranges_off),
        reinterpret_cast<unsigned*>(bin + lay.bktcnt_off),
        reinterpret_cast<unsigned*>(bin + lay.bktlist_off)));
    return 0;
}

// ------------------------------------------------------------------------------------------
// forward.  Warp-autonomous: every warp streams the tile's depth-sorted list by itself, 32
// entries at a time (one per lane, entries three chunks ahead, records one), keeps those whose
// reach mask names its own 8x4 pixel block (ballot) and appends their records, compacted, to a
// warp-private ring in shared memory.  The ring is consumed FWD_U pairs at a time by a two-stage
// software pipeline: EVAL computes the alphas of the next FWD_U pairs (independent chains:
// falloff, exp, clamp) while COMPOSITE folds the previous FWD_U into the pixel state.  COMPOSITE
// is branch-free and its only serial dependency per pair is one multiply (T), one predicate
// (done) and the colour FMAs.  No CTA barrier anywhere; a warp leaves as soon as its 32 pixels
// have saturated.
//
// By-product for the backward: only ~10 % of a tile's entries reach any one pixel block, and the
// scan that finds them is a third of this kernel's instructions.  The warp therefore writes the
// entries it keeps -- (Gaussian id, position in the tile's list) -- to the block's own list
// (plane `block` of 8, at the tile's start in the sorted list: no allocation, planes as long as
// the pair list, mostly untouched), records per pixel the last contributor's index in THAT list,
// and files one backward work item (tile, block, entries to walk).  The backward then streams
// exactly the entries it needs, longest items first, and empty pixel blocks cost it nothing.
//
// EVAL works on TWO ring entries per instruction (packed binary32 pairs, FFMA2 / FMUL2 /
// FADD2 of sm_100: the kernel is issue-bound, and the falloff + exp chain is 19 of its ~33
// floating-point operations per pair).  The ring therefore keeps the geometric part of two
// consecutive entries interleaved -- (x0, x1, y0, y1), (A0, A1, B0, B1), (C0, C1, o0, o1) --
// so one broadcast LDS.128 delivers an operand pair in an aligned register pair; the colour
// part stays one float4 per entry for COMPOSITE.  Lane by lane the operations and their
// order are those of the scalar formulation: results are unchanged, bit for bit.
// ------------------------------------------------------------------------------------------
#ifdef SGS_BLEND_TRACE
// measurement build only (tools/blend_trace.py): one record per work item of the forward blend
struct TraceRec { unsigned long long t0, t1; unsigned smid, tile, warp, relevant, len, pad; };
__device__ TraceRec g_trace[1 << 16];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#endif
constexpr int RING_SLOTS = 64;     // >= 32 + FWD_U
static_assert(FWD_U == 4 && BWD_U == 4, "entries are evaluated in packed pairs; the four list positions of a "
              "batch come from one LDS.128");

struct RingGeo {                    // two consecutive ring entries per element
    float4 xy[RING_SLOTS / 2];      // x0, x1, y0, y1
    float4 ab[RING_SLOTS / 2];      // -a/2 (x2), -b (x2)
    float4 co[RING_SLOTS / 2];      // -c/2 (x2), opacity (x2)
};
__device__ __forceinline__ void ring_clear(RingGeo& g, int lane) {      // ring slots always hold finite values
    g.xy[lane] = g.ab[lane] = g.co[lane] = make_float4(0, 0, 0, 0);
}
__device__ __forceinline__ void ring_put(RingGeo& g, unsigned slot, const float4 q0, const float4 q1) {
    const unsigned ps = slot >> 1, e = slot & 1u;
    float* xy = reinterpret_cast<float*>(&g.xy[ps]);
    float* ab = reinterpret_cast<float*>(&g.ab[ps]);
    float* co = reinterpret_cast<float*>(&g.co[ps]);
    xy[e] = q0.x; xy[2 + e] = q0.y;
    ab[e] = q0.z; ab[2 + e] = q0.w;
    co[e] = q1.x; co[2 + e] = q1.y;
}
// exponent ("power"), falloff G = exp(min(power, 0)) and opacity * G of one packed pair of entries
struct PairEval {
    float2 power, G, oG;
};
__device__ __forceinline__ PairEval ring_eval(const RingGeo& g, unsigned ps, const float2 npx, const float2 npy) {
    const float4 a = g.xy[ps], b = g.ab[ps], c = g.co[ps];
    const float2 dx = fadd2(make_float2(a.x, a.y), npx), dy = fadd2(make_float2(a.z, a.w), npy);
    const float2 uu = fmul2(make_float2(b.x, b.y), dx), vv = fmul2(make_float2(c.x, c.y), dy);
    const float2 ww = fmul2(make_float2(b.z, b.w), dx);
    PairEval r;
    r.power = ffma2(ww, dy, ffma2(vv, dy, fmul2(uu, dx)));
    // (no min(power, 0): an entry with power > 0 is rejected by its caller whatever G comes out)
    r.G = expneg2(r.power);
    r.oG = fmul2(make_float2(c.z, c.w), r.G);
    return r;
}

struct FwdBatch {
    float al[FWD_U], om[FWD_U];
    float4 c[FWD_U];                // r, g, b, depth (AUX)
    unsigned pos[FWD_U];            // index in the block's list + 1, 0 = rejected
};

// AUX: the caller wants the depth image too (GaussianRasterizer.forward_aux): the fourth float of
// the ring's colour entry is the depth, accumulated in the spare half of a packed FMA.
struct FwdWarpSmem {
    RingGeo geo;
    float4 c[RING_SLOTS];
    unsigned pos[RING_SLOTS];        // position in the tile's list + 1 (forward-only frames, which keep no block lists)
};

#ifndef SGS_FWD_CTAS
#define SGS_FWD_CTAS 5             // resident CTAs per SM the register budget is cut for (102 registers at 128 threads)
#endif
// FOR_BWD: leave the block lists, n_blk and the work items for the backward (a forward-only frame --
// animation, evaluation -- skips them and tracks the tile-list position of the last contributor itself).
template <bool AUX, bool FOR_BWD>
__global__ void __launch_bounds__(FWD_WPC * 32, SGS_FWD_CTAS)
blend_fwd_kernel(const uint2* __restrict__ ranges, const unsigned* __restrict__ bucket_count,
                 const unsigned* __restrict__ bucket_list, int tiles,
                 const unsigned* __restrict__ point_list, const float4* __restrict__ rec,
                 uint2* __restrict__ blk_list, size_t list_plane,
                 unsigned* __restrict__ item_count, uint4* __restrict__ item_list,
                 const float* __restrict__ bg, int W, int H, int gx_tiles,
                 float* __restrict__ out_color, float* __restrict__ final_T,
                 unsigned* __restrict__ n_contrib, unsigned* __restrict__ n_blk, float* __restrict__ out_alpha,
                 float* __restrict__ out_depth) {
    __shared__ __align__(16) FwdWarpSmem s_fwd[FWD_WPC];

    const int tid = threadIdx.x, lane = tid & 31, cwarp = tid >> 5;
    FwdWarpSmem& sm = s_fwd[cwarp];
    ring_clear(sm.geo, lane);
    sm.c[lane] = sm.c[lane + 32] = make_float4(0, 0, 0, 0);
    pdl_sync();
  for_each_tile_item<FWD_WPC>(bucket_count, bucket_list, tiles, [&](const unsigned tile, const int warp) {
    const int tile_x = (int)(tile % (unsigned)gx_tiles), tile_y = (int)(tile / (unsigned)gx_tiles);
    int lx, ly;
    pixel_of_thread(warp, lane, lx, ly);
    const int px = tile_x * TILE + lx, py = tile_y * TILE + ly;
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    const float2 npx = splat2(-(float)px), npy = splat2(-(float)py);
#ifdef SGS_BLEND_TRACE
    const unsigned long long trace_t0 = gtimer();
#endif

    bool done = !inside;
    float T = 1.0f;
    float2 C01 = splat2(0.0f), C2D = splat2(0.0f);     // (red, green), (blue, depth)
    unsigned last = 0;
    unsigned head = 0, tail = 0;       // ring: consumed / produced pair counts (warp-uniform)

    FwdBatch bat_a, bat_b;             // one waits to be composited (starts empty), one is being evaluated
    bool cur_is_b = false;             // which of the two is waiting (warp-uniform)
#pragma unroll
    for (int u = 0; u < FWD_U; u++) {
        bat_a.al[u] = 0.0f; bat_a.om[u] = 1.0f; bat_a.c[u] = make_float4(0, 0, 0, 0); bat_a.pos[u] = 0;
    }

    // alphas of ring entries [base, base + FWD_U) (base is a multiple of FWD_U); with LIMIT,
    // entries at or beyond `limit` are empty (only the drain meets a partial batch).  A rejected
    // entry gets alpha 0 (1 - alpha = 1 exactly) and index 0.
    auto eval = [&](FwdBatch& e, unsigned base, unsigned limit, auto LIMIT) {
        const unsigned s0 = base & (RING_SLOTS - 1);
        unsigned pos4[4] = {0u, 0u, 0u, 0u};
        if (!FOR_BWD) {
            const uint4 pp = *reinterpret_cast<const uint4*>(&sm.pos[s0]);
            pos4[0] = pp.x; pos4[1] = pp.y; pos4[2] = pp.z; pos4[3] = pp.w;
        }
#pragma unroll
        for (int h = 0; h < FWD_U / 2; h++) {
            const PairEval pe = ring_eval(sm.geo, (s0 >> 1) + h, npx, npy);
            const float pw[2] = {pe.power.x, pe.power.y}, og[2] = {pe.oG.x, pe.oG.y};
            float al2[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int u = 2 * h + j;
                const float al = fminf(0.99f, og[j]);
                bool ok = !(pw[j] > 0.0f) && !(al < 1.0f / 255.0f);
                if (decltype(LIMIT)::value) ok = ok && (base + u < limit);
                al2[j] = ok ? al : 0.0f;
                e.al[u] = al2[j];
                e.c[u] = sm.c[s0 + u];
                e.pos[u] = ok ? (FOR_BWD ? base + u + 1u : pos4[u]) : 0u;
            }
            const float2 om = ffma2(make_float2(al2[0], al2[1]), splat2(-1.0f), splat2(1.0f));    // 1 - alpha
            e.om[2 * h] = om.x; e.om[2 * h + 1] = om.y;
        }
    };
    // fold a batch into the pixel state, in order ([upstream] renderCUDA: test_T = T (1 - alpha);
    // test_T < 1e-4 -> done, the entry is NOT applied; else C += c alpha T, T = test_T, the entry
    // becomes the last contributor).  Branch-free: T freezes when the pixel saturates (so it ends
    // as the reported final_T), the weight of every later entry is 0.  A rejected entry has
    // alpha 0: test_T = T exactly, and T >= 1e-4 as long as the pixel is live, so it needs no test
    // of its own.  Colour and depth accumulate in two packed FMAs (C0, C1), (C2, depth).
    auto composite = [&](const FwdBatch& e) {
#pragma unroll
        for (int u = 0; u < FWD_U; u++) {
            const float test_T = __fmul_rn(T, e.om[u]);
            done = done || test_T < 0.0001f;
            const float wgt = done ? 0.0f : __fmul_rn(e.al[u], T);
            C01 = ffma2(make_float2(e.c[u].x, e.c[u].y), splat2(wgt), C01);
            if (AUX) C2D = ffma2(make_float2(e.c[u].z, e.c[u].w), splat2(wgt), C2D);
            else C2D.x = __fmaf_rn(e.c[u].z, wgt, C2D.x);
            if (!done) last = max(last, e.pos[u]);
            T = done ? T : test_T;
        }
    };
    const std::true_type with_limit;
    const std::false_type no_limit;

    // Staging pipeline, per chunk of 32 list entries (one per lane): the entries (id | mask) run
    // three chunks ahead, the records of the relevant ones one chunk ahead -- the dependent
    // entry -> record gather never sits on the critical path, and a chunk without relevant
    // entries costs a ballot.
    // a list entry = Gaussian id | reach mask << 24: bit (24 + w) says pixel block w can be reached
    const unsigned* pl = point_list + range.x;
    uint2* bl = blk_list + (size_t)warp * list_plane + range.x;        // this block's own list
    const unsigned wbit = 1u << (ID_BITS + warp);
    auto entry_at = [&](int at) { return at + lane < len ? __ldg(pl + at + lane) : 0u; };
    unsigned m0 = entry_at(0), m1 = entry_at(32), m2 = entry_at(64);
    Rec p;
    p.q0 = p.q1 = p.q2 = make_float4(0, 0, 0, 0);
    if (m0 & wbit) p = load_rec(rec, m0 & ID_MASK);
    for (int pos = 0; pos < len; pos += 32) {
        if (__all_sync(0xffffffffu, done)) break;
        const bool rel = (m0 & wbit) != 0u;
        const unsigned bits = __ballot_sync(0xffffffffu, rel);
        __syncwarp();                  // earlier ring reads are complete before slots are reused
        if (rel) {
            const unsigned idx = tail + __popc(bits & lanemask_lt());
            const unsigned slot = idx & (RING_SLOTS - 1);
            ring_put(sm.geo, slot, p.q0, p.q1);
            sm.c[slot] = make_float4(p.q1.w, p.q2.x, p.q2.y, p.q2.z);
            if (FOR_BWD) bl[idx] = make_uint2(m0 & ID_MASK, (unsigned)(pos + lane));
            else sm.pos[slot] = (unsigned)(pos + lane + 1);
        }
        tail += __popc(bits);
        __syncwarp();
        if (m1 & wbit) p = load_rec(rec, m1 & ID_MASK);
        m0 = m1; m1 = m2; m2 = entry_at(pos + 96);
        // the two batch register sets swap roles every step (no register copies)
        while (tail - head >= FWD_U) {
            if (!cur_is_b) { eval(bat_b, head, tail, no_limit); composite(bat_a); }
            else           { eval(bat_a, head, tail, no_limit); composite(bat_b); }
            cur_is_b = !cur_is_b;
            head += FWD_U;
        }
    }
    // drain: the waiting batch, then the (partial) remainder of the ring.  Nothing to do when no
    // entry ever reached this pixel block -- three quarters of the tiles of an avatar frame are
    // empty, and their warps would spend ~400 instructions compositing empty batches.
    if (tail != 0) {
        if (!cur_is_b) { eval(bat_b, head, tail, with_limit); composite(bat_a); composite(bat_b); }
        else           { eval(bat_a, head, tail, with_limit); composite(bat_b); composite(bat_a); }
    }
    // the backward's work item: the entries of this block's list up to the last contributor of any pixel
    const unsigned wlast = FOR_BWD ? __reduce_max_sync(0xffffffffu, last) : 0u;
    if (wlast != 0u && lane == 0) {
        const unsigned bk = length_bucket(wlast);
        const unsigned slot = atomicAdd(&item_count[bk], 1u);
        item_list[(size_t)bk * ((size_t)tiles * 8u) + slot] = make_uint4(tile * 8u + (unsigned)warp, range.x, wlast, 0u);
    }
    __syncwarp();                                // the list entries written by other lanes are visible
    if (inside) {
        final_T[pix] = T;                        // a saturated pixel reports the T it stopped at
        if (FOR_BWD) {
            n_blk[pix] = last;                   // last contributor: index in the block's list + 1 (the backward starts there)
            n_contrib[pix] = last ? bl[last - 1u].y + 1u : 0u;        // ... and its position in the tile's list + 1 ([upstream] n_contrib)
        } else {
            n_contrib[pix] = last;
        }
        out_color[pix] = __fmaf_rn(T, bg[0], C01.x);
        out_color[plane + pix] = __fmaf_rn(T, bg[1], C01.y);
        out_color[2 * plane + pix] = __fmaf_rn(T, bg[2], C2D.x);
        if (out_alpha) out_alpha[pix] = __fsub_rn(1.0f, T);
        if (out_depth) out_depth[pix] = C2D.y;
    }
#ifdef SGS_BLEND_TRACE
    if (lane == 0) {
        TraceRec r; r.t0 = trace_t0; r.t1 = gtimer(); r.smid = smid(); r.tile = tile; r.warp = (unsigned)warp;
        r.relevant = tail; r.len = (unsigned)len; r.pad = head;
        g_trace[(tile * TILE_WARPS + warp) & 0xffff] = r;
    }
#endif
  });
}
#ifdef SGS_BLEND_TRACE
extern "C" int sgs_debug_trace_read(void* host, size_t bytes) {
    return (int)cudaMemcpyFromSymbol(host, g_trace, bytes);
}
#endif

int launch_blend_fwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     char* img, const float* bg, float* out_color, float* out_alpha,
                     float* out_depth, bool for_backward, cudaStream_t stream) {
    const long long blocks = (long long)lay.tiles * (TILE_WARPS / FWD_WPC);
    auto kernel = for_backward ? (out_depth ? blend_fwd_kernel<true, true> : blend_fwd_kernel<false, true>)
                               : (out_depth ? blend_fwd_kernel<true, false> : blend_fwd_kernel<false, false>);
    SGS_CUDA_OK(launch_pdl(kernel, (unsigned)blocks, FWD_WPC * 32, 0, stream,
        reinterpret_cast<const uint2*>(bin + lay.ranges_off),
        reinterpret_cast<const unsigned*>(bin + lay.bktcnt_off),
        reinterpret_cast<const unsigned*>(bin + lay.bktlist_off), lay.tiles, sorted_vals(lay, const_cast<char*>(bin)),
        reinterpret_cast<const float4*>(geom + lay.rec_off),
        reinterpret_cast<uint2*>(const_cast<char*>(bin) + lay.blklist_off), lay.plane_entries,
        reinterpret_cast<unsigned*>(const_cast<char*>(bin) + lay.itemcnt_off),
        reinterpret_cast<uint4*>(const_cast<char*>(bin) + lay.itemlist_off),
        bg, W, H, lay.gx, out_color,
        reinterpret_cast<float*>(img + lay.finalT_off),
        reinterpret_cast<unsigned*>(img + lay.ncontrib_off),
        reinterpret_cast<unsigned*>(img + lay.nblk_off), out_alpha, out_depth));
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward.  Same warp-autonomous streaming and ring queue, back to front, starting at the
// largest contributor count among the warp's 32 pixels.  Two phases per warp:
//  phase 1 (lane = pixel): the order-dependent part.  EVAL computes falloff G, alpha and
//    1/(1-alpha) of the next BWD_U pairs (two entries per packed instruction, like the
//    forward) while SEQ advances the pixel state (T, the colour behind the pair) over the
//    previous BWD_U, branch-free.  Of everything the nine parameter gradients need, only TWO
//    numbers per (pixel, pair) depend on the order: w = G dL/dalpha and d = alpha T.  SEQ parks
//    them in a warp-private [pair][pixel] tile (row stride 36: conflict-free both ways, rows
//    16-byte aligned).
//  phase 2 (lane = pair, every 32 queued pairs): each lane walks the 32 pixels of its pair's
//    row FOUR AT A TIME (LDS.128 + packed FMAs) and accumulates  sum w, w dx, w dy, w dx^2,
//    w dx dy, w dy^2, d dL/dpix_rgb  in registers -- the reduction over pixels costs no
//    shuffles at all -- then folds them into the nine gradients and issues 2 vector + 1
//    scalar reduction for its Gaussian.
// ------------------------------------------------------------------------------------------
constexpr int BWD_ROW = 36;        // padded row of the [pair][pixel] tiles (16-byte aligned rows)
constexpr int BWD_ROWS = 32;       // pair rows per [pair][pixel] tile
static_assert(BWD_ROWS % BWD_U == 0, "a [pair][pixel] tile must hold whole batches");

struct BwdBatch {
    float al[BWD_U], G[BWD_U], inv[BWD_U], b[BWD_U];
    float2 rg[BWD_U];
};

constexpr int ID_SLOTS = 128;       // Gaussian ids by consumption index: must outlive the ring slot (until phase 2)
struct BwdWarpSmem {
    RingGeo geo;
    float4 c[RING_SLOTS];           // r, g, b, list position (bits)
    float w[BWD_ROWS * BWD_ROW];    // G * dL/dalpha   [pair][pixel]
    float d[BWD_ROWS * BWD_ROW];    // alpha * T       [pair][pixel]
    float dp[3][32];                // dL/dpixel rgb of the warp's 32 pixels
    unsigned id[ID_SLOTS];          // Gaussian of each queued pair
};

__global__ void __launch_bounds__(BWD_WPC * 32, SGS_BWD_MINB * (TILE_WARPS / BWD_WPC))
blend_bwd_kernel(const unsigned* __restrict__ bucket_count,
                 const uint4* __restrict__ bucket_list, int tiles,
                 const uint2* __restrict__ blk_list, size_t list_plane, const float4* __restrict__ rec,
                 const float* __restrict__ bg, int W, int H, int gx_tiles,
                 const float* __restrict__ final_T, const unsigned* __restrict__ n_blk,
                 const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    extern __shared__ __align__(16) char s_bwd_raw[];
    const int tid = threadIdx.x, lane = tid & 31, cwarp = tid >> 5;
    BwdWarpSmem& sm = reinterpret_cast<BwdWarpSmem*>(s_bwd_raw)[cwarp];
    ring_clear(sm.geo, lane);
    sm.c[lane] = sm.c[lane + 32] = make_float4(0, 0, 0, 0);
    pdl_sync();
  for_each_item(bucket_count, bucket_list, tiles, [&](const uint4 item) {
    const unsigned tile = item.x >> 3;
    const int warp = (int)(item.x & 7u);
    const int tile_x = (int)(tile % (unsigned)gx_tiles), tile_y = (int)(tile / (unsigned)gx_tiles);
    int lx, ly;
    pixel_of_thread(warp, lane, lx, ly);
    const int px = tile_x * TILE + lx, py = tile_y * TILE + ly;
    const bool inside = px < W && py < H;
    const float2 npx = splat2(-(float)px), npy = splat2(-(float)py);
    const float bx0 = (float)(tile_x * TILE + ((warp & 1) << 3));
    const float by0 = (float)(tile_y * TILE + ((warp >> 1) << 2));
    const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;

    // last contributor of the pixel: index in the block's list + 1
    const unsigned last = inside ? n_blk[pix] : 0u;
    // entries at or beyond the warp's largest index reach none of its pixels
    const int wlast = (int)__reduce_max_sync(0xffffffffu, last);
    if (wlast == 0) return;
    __syncwarp();                              // (a previous item's last reduction has read dp)
    const float T_final = inside ? final_T[pix] : 0.0f;
    float dp0 = 0.0f, dp1 = 0.0f, dp2 = 0.0f;
    if (inside) { dp0 = dL_dpix[pix]; dp1 = dL_dpix[plane + pix]; dp2 = dL_dpix[2 * plane + pix]; }
    sm.dp[0][lane] = dp0; sm.dp[1][lane] = dp1; sm.dp[2][lane] = dp2;
    const float nTf_bg = -T_final * (bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2);
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;

    float T = T_final;
    float2 B01 = splat2(0.0f);                 // colour accumulated behind the current pair (red, green)
    float B2 = 0.0f;                           // ... blue
    unsigned head = 0, tail = 0;               // ring: consumed / produced pair counts
    unsigned row0 = 0;                         // first pair (consumption index) of the open [pair][pixel] tile; multiple of BWD_ROWS

    BwdBatch bat_a, bat_b;             // one waits for SEQ (starts empty), one is being evaluated
    bool cur_is_b = false;
#pragma unroll
    for (int u = 0; u < BWD_U; u++) {
        bat_a.al[u] = 0.0f; bat_a.G[u] = 0.0f; bat_a.inv[u] = 1.0f;
        bat_a.rg[u] = splat2(0.0f); bat_a.b[u] = 0.0f;
    }

    auto eval = [&](BwdBatch& e, unsigned base, unsigned limit, auto LIMIT) {
        const unsigned s0 = base & (RING_SLOTS - 1);
#pragma unroll
        for (int h = 0; h < BWD_U / 2; h++) {
            const PairEval pe = ring_eval(sm.geo, (s0 >> 1) + h, npx, npy);
            const float pw[2] = {pe.power.x, pe.power.y}, og[2] = {pe.oG.x, pe.oG.y}, Gj[2] = {pe.G.x, pe.G.y};
            float al2[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int u = 2 * h + j;
                const float4 c = sm.c[s0 + u];
                const float al = fminf(0.99f, og[j]);
                bool ok = __float_as_uint(c.w) < last && !(pw[j] > 0.0f) && !(al < 1.0f / 255.0f);
                if (decltype(LIMIT)::value) ok = ok && (base + u < limit);
                al2[j] = ok ? al : 0.0f;
                e.al[u] = al2[j];
                e.G[u] = ok ? Gj[j] : 0.0f;
                e.rg[u] = make_float2(c.x, c.y); e.b[u] = c.z;
            }
            const float2 om = ffma2(make_float2(al2[0], al2[1]), splat2(-1.0f), splat2(1.0f));
            e.inv[2 * h] = __fdividef(1.0f, om.x); e.inv[2 * h + 1] = __fdividef(1.0f, om.y);
        }
    };
    // advance the pixel state over a batch whose first pair has consumption index `base`
    auto seq = [&](const BwdBatch& e, unsigned base) {
#pragma unroll
        for (int u = 0; u < BWD_U; u++) {
            T = T * e.inv[u];
            const float dch = e.al[u] * T;
            const float2 t01 = ffma2(B01, splat2(-1.0f), e.rg[u]);      // colour - colour behind
            const float t2 = e.b[u] - B2;
            const float dot = fmaf(t2, dp2, fmaf(t01.y, dp1, t01.x * dp0));
            const float dLda = fmaf(T, dot, nTf_bg * e.inv[u]);
            const unsigned row = (base + u) % BWD_ROWS;
            sm.w[row * BWD_ROW + lane] = e.G[u] * dLda;
            sm.d[row * BWD_ROW + lane] = dch;
            B01 = ffma2(splat2(e.al[u]), t01, B01);       // = al * c + (1 - al) * B
            B2 = fmaf(e.al[u], t2, B2);
        }
    };
    // phase 2 over the `cnt` (<= BWD_ROWS) pair rows of the open tile.  Lane = pair row r; it
    // walks the 32 pixels four at a time (even / odd columns in the two halves of a packed
    // pair) accumulating RAW moments of w about the block origin and re-centres them on the
    // Gaussian after the loop:
    //   sum w dx = ax S0 - Mx,  sum w dx^2 = ax (ax S0 - 2 Mx) + Mxx, ... (dx = ax - kx).
    // With kx = 2k (+1 in the odd half) the moments need only broadcast scalar factors:
    //   sum kx w = sum 2k (we + wo) + sum wo,   sum kx^2 w = sum 4k^2 (we + wo) + 2 sum 2k wo + sum wo.
    auto reduce_rows = [&](unsigned cnt) {
        __syncwarp();
        const unsigned r = lane;
        const bool valid = r < cnt;
        const unsigned id = valid ? sm.id[(row0 + r) & (ID_SLOTS - 1)] : 0u;
        const float4 q0 = __ldg(rec + 4 * (size_t)id);
        const float4 q1 = __ldg(rec + 4 * (size_t)id + 1);
        float S0 = 0, Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0;
        float2 Sr = splat2(0.0f), Sg = splat2(0.0f), Sb = splat2(0.0f);
        const float4* wr = reinterpret_cast<const float4*>(sm.w + r * BWD_ROW);
        const float4* dr = reinterpret_cast<const float4*>(sm.d + r * BWD_ROW);
        const float4* dpr = reinterpret_cast<const float4*>(&sm.dp[0][0]);
#pragma unroll 1
        for (int yy = 0; yy < 4; yy++) {                      // one pixel row of the 8x4 block per trip
            const float ky = (float)yy;
            float2 R0 = splat2(0.0f), Rx = splat2(0.0f), Rxx = splat2(0.0f);      // this row's moments in x (even, odd columns)
#pragma unroll
            for (int k4 = 0; k4 < 2; k4++) {
                const float4 w4 = wr[yy * 2 + k4], d4 = dr[yy * 2 + k4];
                const float4 pr = dpr[yy * 2 + k4], pg = dpr[8 + yy * 2 + k4], pb = dpr[16 + yy * 2 + k4];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int kp = 2 * k4 + j;                // columns 2 kp, 2 kp + 1
                    const float2 w = j ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
                    const float2 d = j ? make_float2(d4.z, d4.w) : make_float2(d4.x, d4.y);
                    R0 = fadd2(R0, w);
                    if (kp != 0) {
                        Rx = ffma2(w, splat2((float)(2 * kp)), Rx);
                        Rxx = ffma2(w, splat2((float)(4 * kp * kp)), Rxx);
                    }
                    Sr = ffma2(d, j ? make_float2(pr.z, pr.w) : make_float2(pr.x, pr.y), Sr);
                    Sg = ffma2(d, j ? make_float2(pg.z, pg.w) : make_float2(pg.x, pg.y), Sg);
                    Sb = ffma2(d, j ? make_float2(pb.z, pb.w) : make_float2(pb.x, pb.y), Sb);
                }
            }
            const float r0 = R0.x + R0.y, rx = (Rx.x + Rx.y) + R0.y, rxx = (Rxx.x + Rxx.y) + fmaf(2.0f, Rx.y, R0.y);
            S0 += r0; Mx += rx; Mxx += rxx;
            My = fmaf(r0, ky, My); Myy = fmaf(r0, ky * ky, Myy); Mxy = fmaf(rx, ky, Mxy);
        }
        const float ax = q0.x - bx0, ay = q0.y - by0;
        const float Sx = fmaf(ax, S0, -Mx), Sy = fmaf(ay, S0, -My);
        const float Sxx = fmaf(ax, fmaf(ax, S0, -2.0f * Mx), Mxx);
        const float Syy = fmaf(ay, fmaf(ay, S0, -2.0f * My), Myy);
        const float Sxy = fmaf(ax, fmaf(ay, S0, -My), fmaf(-ay, Mx, Mxy));
        if (valid) {
            // conic entries: a = -2*q0.z, b = -q0.w, c = -2*q1.x
            const float ca = -2.0f * q0.z, cb = -q0.w, cc = -2.0f * q1.x, o = q1.y;
            float* dst = acc + (size_t)id * ACC_FLOATS;
            // accumulator slots 0..8: mean2D.x, .y, conic a, b, c, opacity, r, g, b
            red_add_f4(dst, -o * ddelx_dx * (ca * Sx + cb * Sy), -o * ddely_dy * (cc * Sy + cb * Sx),
                       -0.5f * o * Sxx, -0.5f * o * Sxy);
            red_add_f4(dst + 4, -0.5f * o * Syy, S0, Sr.x + Sr.y, Sg.x + Sg.y);
            atomicAdd(dst + 8, Sb.x + Sb.y);
        }
        __syncwarp();
    };
    const std::true_type with_limit;
    const std::false_type no_limit;

    // The block's list back to front: round r covers indices top - lane, top = wlast - 1 - 32 r, so
    // lane order is back-to-front order.  Entries two rounds ahead, their records one round
    // ahead, in registers (like the forward).
    const uint2* bl = blk_list + (size_t)warp * list_plane + item.y;
    auto entry_at = [&](int top) { return top - lane >= 0 ? __ldg(bl + top - lane).x : 0u; };
    unsigned e1 = entry_at(wlast - 33), e2 = entry_at(wlast - 65);
    Rec p;
    unsigned pid = entry_at(wlast - 1);
    p.q0 = p.q1 = p.q2 = make_float4(0, 0, 0, 0);
    if (wlast - 1 - lane >= 0) p = load_rec(rec, pid);
    // `cur` starts as an empty batch at consumption index -BWD_U: its SEQ writes zeros into
    // rows that real pairs overwrite before any reduction reads them.
    unsigned pend = 0u - BWD_U;
    for (int top = wlast - 1; top >= 0; top -= 32) {
        const int nv = min(32, top + 1);
        __syncwarp();
        if (lane < nv) {
            const unsigned idx = tail + lane;              // consumption index of this entry
            const unsigned slot = idx & (RING_SLOTS - 1);
            ring_put(sm.geo, slot, p.q0, p.q1);
            sm.c[slot] = make_float4(p.q1.w, p.q2.x, p.q2.y, __uint_as_float((unsigned)(top - lane)));
            sm.id[idx & (ID_SLOTS - 1)] = pid;
        }
        tail += nv;
        __syncwarp();
        if (top - 32 - lane >= 0) {
            pid = e1;
            p = load_rec(rec, pid);
        }
        e1 = e2; e2 = entry_at(top - 96);
        // the two batch register sets swap roles every step (no register copies)
        while (tail - head >= BWD_U) {
            if (!cur_is_b) { eval(bat_b, head, tail, no_limit); seq(bat_a, pend); }
            else           { eval(bat_a, head, tail, no_limit); seq(bat_b, pend); }
            cur_is_b = !cur_is_b;
            if (pend + BWD_U - row0 == BWD_ROWS) {       // the open tile is full
                reduce_rows(BWD_ROWS);
                row0 += BWD_ROWS;
            }
            pend = head;
            head += BWD_U;
        }
    }
    // drain: the waiting batch, the partial remainder of the ring, the partial tile
    if (!cur_is_b) { eval(bat_b, head, tail, with_limit); seq(bat_a, pend); }
    else           { eval(bat_a, head, tail, with_limit); seq(bat_b, pend); }
    if (pend + BWD_U - row0 == BWD_ROWS) {
        reduce_rows(BWD_ROWS);
        row0 += BWD_ROWS;
    }
    if (tail > head) {
        if (!cur_is_b) seq(bat_b, head); else seq(bat_a, head);
    }
    if (tail > row0) reduce_rows(tail - row0);
  });
}

int launch_blend_bwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     const char* img, const float* bg, const float* dL_dpix, float* acc,
                     cudaStream_t stream) {
    const size_t smem = sizeof(BwdWarpSmem) * BWD_WPC;
    SGS_CUDA_OK(set_max_smem(blend_bwd_kernel, smem));
    SGS_CUDA_OK(launch_pdl(blend_bwd_kernel, blend_grid<BWD_WPC>(lay.tiles), BWD_WPC * 32, smem, stream,
        reinterpret_cast<const unsigned*>(bin + lay.itemcnt_off),
        reinterpret_cast<const uint4*>(bin + lay.itemlist_off), lay.tiles,
        reinterpret_cast<const uint2*>(bin + lay.blklist_off), lay.plane_entries,
        reinterpret_cast<const float4*>(geom + lay.rec_off), bg, W, H, lay.gx,
        reinterpret_cast<const float*>(img + lay.finalT_off),
        reinterpret_cast<const unsigned*>(img + lay.nblk_off), dL_dpix, acc));
    return 0;
}

}  // namespace sgs

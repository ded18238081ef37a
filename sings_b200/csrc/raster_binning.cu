// raster_binning.cu -- the sorted (tile|depth, Gaussian id) pair list by count / scan / scatter.
//
// Replaces, with identical results, the reference rasterizer's
//   [upstream] cub::DeviceScan::InclusiveSum + rasterizer_impl.cu duplicateWithKeys +
//   the tile-id digits of cub::DeviceRadixSort::SortPairs + identifyTileRanges
//   (K4..K7 of section 2.4),
// reached from /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// With the Gaussians already in depth order (launch_depth_sort: the four depth digits are sorted
// once per Gaussian, N items), the sorted pair list is a STABLE counting sort of the pairs by
// tile id: tile t's segment holds, in depth order, the Gaussians whose rectangle covers t.  A
// stable sort has one answer, so the list is bit-identical to sorting all pairs on all 45 bits.
//   bin_count  : chunk c = BIN_GAUSS consecutive depth-ordered Gaussians -> counts[c][t] (u16)
//   bin_scan   : per tile, exclusive prefix of the counts over the chunks -> base[c][t], total[t];
//                its last CTA: tile starts = exclusive scan of total[] -> tile ranges, tiles
//                bucketed by list length, pair count
//   bin_scatter: every pair of the chunk goes to start[t] + base[c][t] + its rank inside the chunk
//   pair_masks : one reach-mask byte per pair (which 8x4 pixel blocks of the tile the Gaussian's
//                footprint can touch); fused into the scatter it ran at 43 % lane efficiency and
//                tripled that kernel's instruction count (profiles/README.md, round 2).
// No radix pass over the pairs, no search for the ranges: the pair list is written once.
//
// Order inside a chunk.  Counting is order-free (shared-memory atomics).  Placement is not: a
// warp owns 128 consecutive Gaussians and takes them in steps that follow depth order, lanes =
// the tiles of the step's rectangles (distinct, so plain read-modify-write of the warp's
// private per-tile counters): ranks are stable by construction.  Everything a step
// needs is staged in shared memory up front -- the step loop touches no global memory except
// its stores (the first version of this scheme fetched rectangles inside the loop and spent
// its time on dependent L2 round trips: 33 us per pass; profiles/README.md, round 1).
#include "common.cuh"
#include "kernels.h"

namespace sgs {

constexpr int BIN_WARPS = BIN_THREADS / 32;
constexpr int BIN_WARP_GAUSS = BIN_GAUSS / BIN_WARPS;      // 128
constexpr int BIN_PER_LANE = BIN_WARP_GAUSS / 32;          // 4
constexpr int SCAN_THREADS = 512;
static_assert(BIN_WARP_GAUSS % 32 == 0, "a warp stages whole rows of 32 Gaussians");

struct BinArgs {
    int P, tiles, tp, gx, ctas;   // tp = row stride of the count / base matrices (tiles padded to 64)
    const unsigned* nkeys[2];     // depth-sorted keys are in buffer depth_sort_parity(varbits)
    const unsigned* nvals[2];
    const unsigned* varbits;
    const uint2* rects;
    unsigned short* counts;       // [ctas][tp]
    unsigned* base;               // [ctas][tp]
    unsigned* total;              // [tp]
    unsigned* start;              // [tp] first pair of every tile (exclusive scan of total)
    const float4* rec;            // geometry records (reach masks)
    unsigned char* masks;         // (L_cap) reach mask per pair (output)
    int* counters;
    int* host_counters;           // mapped host memory for {num_rendered, overflow}, or null
    unsigned long long* keys;     // sorted list (output)
    unsigned* vals;
    uint2* ranges;                // [tiles]
    unsigned* bucket_count;       // [32] (zeroed)
    unsigned* bucket_list;        // [32][tiles]
    long long L_cap;
};

// ---- count: per-chunk pair counts per tile, order-free ----
__global__ void __launch_bounds__(BIN_THREADS) bin_count_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned s_cnt[];            // [tp]
    for (int i = threadIdx.x; i < a.tp / 4; i += BIN_THREADS)
        reinterpret_cast<uint4*>(s_cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
    pdl_sync();
    __syncthreads();
    const int par = depth_sort_parity(a.varbits, DEPTH_PASSES);
    const unsigned* __restrict__ nv = par ? a.nvals[1] : a.nvals[0];
    const int chunk = blockIdx.x;
    uint2 r[BIN_PER_LANE];
#pragma unroll
    for (int i = 0; i < BIN_PER_LANE; i++) {
        const int pos = chunk * BIN_GAUSS + i * BIN_THREADS + threadIdx.x;
        r[i] = pos < a.P ? __ldg(a.rects + nv[pos]) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < BIN_PER_LANE; i++) {
        const unsigned x0 = r[i].x & 0xffffu, y0 = r[i].x >> 16, w = r[i].y & 0xffffu, h = r[i].y >> 16;
        for (unsigned ty = 0; ty < h; ty++) {
            const unsigned row = (y0 + ty) * (unsigned)a.gx + x0;
            for (unsigned tx = 0; tx < w; tx++) atomicAdd(&s_cnt[row + tx], 1u);
        }
    }
    __syncthreads();
    // the chunk's row of the count matrix: two tiles per 32-bit word (counts <= BIN_GAUSS)
    unsigned* out = reinterpret_cast<unsigned*>(a.counts + (size_t)chunk * a.tp);
    for (int i = threadIdx.x; i < a.tp / 2; i += BIN_THREADS) out[i] = s_cnt[2 * i] | (s_cnt[2 * i + 1] << 16);
}

// ---- scan: prefix over the chunks, per tile.  A CTA takes a slab of 64 tiles (lane = two
// tiles = one 32-bit word of a count row); its 16 warps split the chunk rows into contiguous
// groups: sum the group (independent loads, all in flight), exchange the partial sums through
// shared memory, then walk the group again writing the prefixes.  The LAST CTA to finish then
// turns the per-tile totals into tile starts (exclusive scan of a few thousand values, once per
// frame), tile ranges, the length buckets of the blend kernels and the pair count. ----
__global__ void __launch_bounds__(SCAN_THREADS) bin_scan_kernel(BinArgs a) {
    __shared__ uint2 s_part[SCAN_THREADS / 32][32];
    __shared__ unsigned s_wsum[SCAN_THREADS / 32];
    __shared__ unsigned s_bkt[32], s_bkt_base[32];
    __shared__ int s_last;
    constexpr int MAXPER = BIN_MAX_TILES / SCAN_THREADS;       // tiles per thread of the last CTA's scan
    unsigned s_slot[MAXPER], s_bk[MAXPER];                      // (registers: the loops over them are unrolled)
    pdl_sync();
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
    if (tid < 32) s_bkt[tid] = 0;
    const int t2 = blockIdx.x * 32 + lane;                       // tile pair (2 t2, 2 t2 + 1)
    const int rows = (a.ctas + SCAN_THREADS / 32 - 1) / (SCAN_THREADS / 32);
    const int r0 = min(a.ctas, grp * rows), r1 = min(a.ctas, r0 + rows);
    const unsigned* __restrict__ cn = reinterpret_cast<const unsigned*>(a.counts) + t2;
    uint2* __restrict__ bs = reinterpret_cast<uint2*>(a.base) + t2;
    const size_t stride = (size_t)a.tp / 2;
    uint2 sum = make_uint2(0u, 0u);
#pragma unroll 8
    for (int c = r0; c < r1; c++) {
        const unsigned v = __ldg(cn + (size_t)c * stride);
        sum.x += v & 0xffffu; sum.y += v >> 16;
    }
    s_part[grp][lane] = sum;
    __syncthreads();
    uint2 run = make_uint2(0u, 0u);
#pragma unroll
    for (int g = 0; g < SCAN_THREADS / 32; g++)
        if (g < grp) { run.x += s_part[g][lane].x; run.y += s_part[g][lane].y; }
    int c = r0;
    for (; c + 8 <= r1; c += 8) {               // loads first, then the dependent prefix and the stores
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(cn + (size_t)(c + k) * stride);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            bs[(size_t)(c + k) * stride] = run;
            run.x += v[k] & 0xffffu; run.y += v[k] >> 16;
        }
    }
    for (; c < r1; c++) {
        const unsigned v = __ldg(cn + (size_t)c * stride);
        bs[(size_t)c * stride] = run;
        run.x += v & 0xffffu; run.y += v >> 16;
    }
    if (grp == SCAN_THREADS / 32 - 1)           // the last group ends on the grand total
        reinterpret_cast<uint2*>(a.total)[t2] = run;
    // ---- last CTA: tile starts, ranges, buckets, pair count ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&a.counters[CNT_RANGES_DONE], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // `per` consecutive tiles per thread (tp % 512 == 0, per % 4 == 0 ... or 1, 2), one block scan
    const int per = a.tp / SCAN_THREADS;
    const unsigned cap = (unsigned)min(a.L_cap, (long long)0xffffffffll);
    const int tb = tid * per;
    unsigned mine = 0;
    for (int k = 0; k < per; k++) mine += ld_relaxed_u32(a.total + tb + k);      // written by other CTAs of this launch
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    if (lane == 31) s_wsum[grp] = incl;
    __syncthreads();
    unsigned off = 0, running = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const unsigned v = s_wsum[w];
        if (w < grp) off += v;
        running += v;
    }
    unsigned start = off + incl - mine;
#pragma unroll
    for (int k = 0; k < MAXPER; k++) {
        if (k >= per) break;
        const int t = tb + k;
        const unsigned n = ld_relaxed_u32(a.total + t);
        a.start[t] = start;
        if (t < a.tiles) {
            const unsigned lo = min(start, cap), hi = min(start + n, cap);
            a.ranges[t] = hi > lo ? make_uint2(lo, hi) : make_uint2(0u, 0u);
            // file the tile by list length (bucket = bit length of the count)
            s_bk[k] = (unsigned)(32 - __clz(hi - lo)) % 32u;
        } else {
            s_bk[k] = 0xffffffffu;
        }
        start += n;
    }
    // slots inside the buckets: one shared atomic per warp and distinct bucket (three quarters of an
    // avatar frame's tiles are empty -- thousands of same-address atomics otherwise)
#pragma unroll
    for (int k = 0; k < MAXPER; k++) {
        if (k >= per) break;
        const unsigned bk = s_bk[k];
        const unsigned peers = __match_any_sync(0xffffffffu, bk);
        const int leader = __ffs(peers) - 1;
        unsigned slot = 0;
        if (lane == leader && bk != 0xffffffffu) slot = atomicAdd(&s_bkt[bk], (unsigned)__popc(peers));
        s_slot[k] = __shfl_sync(0xffffffffu, slot, leader) + (unsigned)__popc(peers & lanemask_lt());
    }
    __syncthreads();
    // bucket bases in the global lists (this CTA is the only writer), then the tiles
    if (tid < 32) s_bkt_base[tid] = atomicAdd(&a.bucket_count[tid], s_bkt[tid]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXPER; k++) {
        if (k >= per) break;
        const int t = tb + k;
        if (s_bk[k] != 0xffffffffu) a.bucket_list[(size_t)s_bk[k] * a.tiles + s_bkt_base[s_bk[k]] + s_slot[k]] = (unsigned)t;
    }
    if (tid == 0) {
        const unsigned long long total = running;
        a.counters[CNT_NUM_RENDERED] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
        if (total > (unsigned long long)a.L_cap) a.counters[CNT_OVERFLOW] = 1;
        if (a.host_counters) {
            a.host_counters[0] = total > 0x7fffffffull ? 0x7fffffff : (int)total;
            a.host_counters[1] = total > (unsigned long long)a.L_cap ? 1 : 0;
            __threadfence_system();
        }
    }
}

// ---- scatter ----
__global__ void __launch_bounds__(BIN_THREADS) bin_scatter_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned char s_bin[];
    unsigned* s_base = reinterpret_cast<unsigned*>(s_bin);                                        // [tp]
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_bin + (size_t)a.tp * 4);          // [BIN_WARPS][tp]
    uint4* s_info = reinterpret_cast<uint4*>(s_bin + (size_t)a.tp * 4 + (size_t)BIN_WARPS * a.tp * 2);   // [BIN_WARPS][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x;
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        for (int i = tid; i < BIN_WARPS * a.tp * 2 / 16; i += BIN_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_sync();
    // ---- stage this warp's 128 Gaussians: word 0 = x0 | y0 << 8 | w << 16 | h << 24, word 1 =
    // floor(2^32 / w) + 1 (tile index -> row by one multiply-high), word 2 = id, word 3 = depth bits
    const int par = depth_sort_parity(a.varbits, DEPTH_PASSES);
    const unsigned* __restrict__ nk = par ? a.nkeys[1] : a.nkeys[0];
    const unsigned* __restrict__ nv = par ? a.nvals[1] : a.nvals[0];
    uint4* const info = s_info + warp * BIN_WARP_GAUSS;
    uint2 rr[BIN_PER_LANE];
    unsigned gid[BIN_PER_LANE], dk[BIN_PER_LANE];
#pragma unroll
    for (int i = 0; i < BIN_PER_LANE; i++) {
        const int pos = chunk * BIN_GAUSS + warp * BIN_WARP_GAUSS + i * 32 + lane;
        gid[i] = 0; dk[i] = 0;
        if (pos < a.P) { gid[i] = nv[pos]; dk[i] = nk[pos]; }
    }
#pragma unroll
    for (int i = 0; i < BIN_PER_LANE; i++) {
        const int pos = chunk * BIN_GAUSS + warp * BIN_WARP_GAUSS + i * 32 + lane;
        rr[i] = pos < a.P ? __ldg(a.rects + gid[i]) : make_uint2(0u, 0u);
    }
    // ---- where this chunk's pairs of each tile begin: tile start + prefix over the earlier chunks ----
    {
        const unsigned* __restrict__ brow = a.base + (size_t)chunk * a.tp;
        for (int t = tid * 4; t < a.tp; t += BIN_THREADS * 4) {
            const uint4 s = __ldg(reinterpret_cast<const uint4*>(a.start + t));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(brow + t));
            *reinterpret_cast<uint4*>(s_base + t) = make_uint4(s.x + b.x, s.y + b.y, s.z + b.z, s.w + b.w);
        }
    }
    __syncthreads();                                             // the counters are zero
    // ---- pass A: this warp's pair counts per tile (order-free; two tiles per 32-bit word) ----
    unsigned* const cnt32 = reinterpret_cast<unsigned*>(s_cnt + (size_t)warp * a.tp);
    unsigned warp_pairs = 0;
#pragma unroll
    for (int i = 0; i < BIN_PER_LANE; i++) {
        const unsigned x0 = rr[i].x & 0xffffu, y0 = rr[i].x >> 16, w = rr[i].y & 0xffffu, h = rr[i].y >> 16;
        warp_pairs += w * h;
        info[i * 32 + lane] = make_uint4(x0 | (y0 << 8) | (w << 16) | (h << 24), w ? 0xffffffffu / w + 1u : 0u, gid[i], dk[i]);
        for (unsigned ty = 0; ty < h; ty++) {
            const unsigned row = (y0 + ty) * (unsigned)a.gx + x0;
            for (unsigned tx = 0; tx < w; tx++) {
                const unsigned t = row + tx;
                atomicAdd(&cnt32[t >> 1], 1u << (16 * (t & 1u)));
            }
        }
    }
    warp_pairs = __reduce_add_sync(0xffffffffu, warp_pairs);
    __syncthreads();
    // exclusive over the warps, per tile (16-bit halves: no carry, a chunk has <= BIN_GAUSS pairs per tile)
    {
        unsigned* w32 = reinterpret_cast<unsigned*>(s_cnt);
        const int half = a.tp / 2;
        for (int i = tid; i < half; i += BIN_THREADS) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < BIN_WARPS; w++) {
                const unsigned v = w32[(size_t)w * half + i];
                w32[(size_t)w * half + i] = run;
                run += v;
            }
        }
    }
    __syncthreads();
    if (warp_pairs == 0) return;                                 // culled Gaussians sort first: whole warps of them
    // ---- pass B: place.  Steps of FOUR consecutive Gaussians; the step's pairs, in (Gaussian,
    // tile) order, are dealt to the lanes one each (usually one trip: ~22 pairs per step).  Tiles
    // of one Gaussian are distinct; across the four they are distinct when the rectangles are
    // pairwise disjoint (the usual case: depth neighbours are spatially unrelated) -- then every
    // lane owns its tile for the trip and the running per-(warp, tile) counter is a plain
    // read-modify-write.  Otherwise equal tiles are ranked in lane (= depth) order with match_any
    // and the counter advances by the group size.  Trips and steps follow depth order.
    unsigned short* const cnt = s_cnt + (size_t)warp * a.tp;
    const int g4 = lane & 3;
    for (int q = 0; q < BIN_WARP_GAUSS; q += 4) {
        const uint4 fm = info[q + g4];                           // lanes 0..3 hold the step's four Gaussians
        const unsigned xm = fm.x & 0xffu, ym = (fm.x >> 8) & 0xffu, wm = (fm.x >> 16) & 0xffu, hm = fm.x >> 24;
        const unsigned nm = wm * hm;
        const unsigned n0 = __shfl_sync(0xffffffffu, nm, 0), n1 = __shfl_sync(0xffffffffu, nm, 1);
        const unsigned n2 = __shfl_sync(0xffffffffu, nm, 2), n3 = __shfl_sync(0xffffffffu, nm, 3);
        const unsigned p1 = n0, p2 = p1 + n1, p3 = p2 + n2, tot = p3 + n3;
        if (tot == 0) continue;
        bool clash = false;                                      // lane g against g+1 and g+2 (mod 4): all six pairs
#pragma unroll
        for (int d = 1; d <= 2; d++) {
            const unsigned o = __shfl_sync(0xffffffffu, fm.x, (lane + d) & 3);
            const unsigned ox = o & 0xffu, oy = (o >> 8) & 0xffu, ow = (o >> 16) & 0xffu, oh = o >> 24;
            clash = clash || (nm != 0 && ow * oh != 0 && xm < ox + ow && ox < xm + wm && ym < oy + oh && oy < ym + hm);
        }
        const bool any_clash = __any_sync(0xffffffffu, clash);
        for (unsigned e0 = 0; e0 < tot; e0 += 32) {
            const unsigned e = e0 + lane;
            const bool active = e < tot;
            const unsigned g = (unsigned)(e >= p1) + (unsigned)(e >= p2) + (unsigned)(e >= p3);
            const unsigned k = e - (g == 0 ? 0u : g == 1 ? p1 : g == 2 ? p2 : p3);
            const uint4 f = info[q + g];
            const unsigned w = (f.x >> 16) & 0xffu;
            const unsigned ty = w == 1u ? k : __umulhi(k, f.y), tx = k - ty * w;      // (2^32 / 1 does not fit the multiplier)
            const unsigned t = active ? (((f.x >> 8) & 0xffu) + ty) * (unsigned)a.gx + (f.x & 0xffu) + tx : 0u;
            const unsigned c = cnt[t];
            unsigned rank = 0, group = 1;
            if (any_clash) {
                const unsigned peers = __match_any_sync(0xffffffffu, active ? t : (0x80000000u | (unsigned)lane));
                rank = (unsigned)__popc(peers & lanemask_lt());
                group = (unsigned)__popc(peers);
            }
            __syncwarp();                                        // every lane has read its counter
            if (active) {
                cnt[t] = (unsigned short)(c + group);            // the same value from every lane of a group
                const unsigned long long p = (unsigned long long)s_base[t] + c + rank;
                if (p < (unsigned long long)a.L_cap) {
                    a.keys[p] = ((unsigned long long)t << 32) | f.w;
                    a.vals[p] = f.z;
                }
            }
            __syncwarp();                                        // the next trip / step may hit the same tiles
        }
    }
}

// ---- reach masks: one byte per pair of the sorted list; MASK_PER pairs per thread, every load of
// a thread in flight before the first use ----
constexpr int MASK_THREADS = 256;
constexpr int MASK_PER = 1;
__global__ void __launch_bounds__(MASK_THREADS)
pair_masks_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ point_list,
                  const int* __restrict__ counters, long long n_cap, const float4* __restrict__ rec,
                  int gx_tiles, unsigned char* __restrict__ masks) {
    pdl_sync();
    const long long n = min((long long)counters[CNT_NUM_RENDERED], n_cap);
    const long long i0 = (long long)blockIdx.x * (MASK_THREADS * MASK_PER) + threadIdx.x;
    if (i0 >= n) return;
    unsigned tile[MASK_PER], id[MASK_PER];
#pragma unroll
    for (int k = 0; k < MASK_PER; k++) {
        const long long i = min(i0 + k * MASK_THREADS, n - 1);       // out-of-range slots repeat the last pair (not stored)
        tile[k] = (unsigned)(ldg_u64_pinned(keys + i) >> 32);
        id[k] = ldg_u32_pinned(point_list + i);
    }
    float4 q0[MASK_PER], q1[MASK_PER], q3[MASK_PER];
#pragma unroll
    for (int k = 0; k < MASK_PER; k++) {
        const float4* p = rec + 4 * (size_t)id[k];
        q0[k] = ldg_f4_pinned(p); q1[k] = ldg_f4_pinned(p + 1); q3[k] = ldg_f4_pinned(p + 3);
    }
#pragma unroll
    for (int k = 0; k < MASK_PER; k++) {
        const long long i = i0 + k * MASK_THREADS;
        const float tx = (float)((tile[k] % (unsigned)gx_tiles) * TILE), ty = (float)((tile[k] / (unsigned)gx_tiles) * TILE);
        const unsigned m = reach_mask(q0[k], q1[k], q3[k], tx, ty);
        if (i < n) masks[i] = (unsigned char)m;
    }
}

bool bin_css_supported(const RasterLayout& lay) {
    return lay.tiles <= BIN_MAX_TILES && lay.gx <= 255 && lay.gy <= 255;
}

int launch_bin_css(int P, const RasterLayout& lay, long long L_cap, const char* geom, char* bin,
                   int* host_counters, cudaStream_t stream, int debug) {
    if (P <= 0) return 0;
    BinArgs a;
    a.P = P; a.tiles = lay.tiles; a.tp = lay.bin_tp; a.gx = lay.gx; a.ctas = lay.bin_ctas;
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
    a.start = reinterpret_cast<unsigned*>(bin + lay.bstart_off);
    a.rec = reinterpret_cast<const float4*>(geom + lay.rec_off);
    a.masks = reinterpret_cast<unsigned char*>(bin + lay.masks_off);
    a.host_counters = host_counters;
    a.keys = reinterpret_cast<unsigned long long*>(bin + (lay.sorted_in_1() ? lay.keys1_off : lay.keys0_off));
    a.vals = reinterpret_cast<unsigned*>(bin + (lay.sorted_in_1() ? lay.vals1_off : lay.vals0_off));
    a.ranges = reinterpret_cast<uint2*>(bin + lay.ranges_off);
    a.bucket_count = reinterpret_cast<unsigned*>(bin + lay.bktcnt_off);
    a.bucket_list = reinterpret_cast<unsigned*>(bin + lay.bktlist_off);
    a.L_cap = L_cap;
    const size_t smem_count = (size_t)lay.bin_tp * 4;
    const size_t smem_scatter = (size_t)lay.bin_tp * 4 + (size_t)BIN_WARPS * lay.bin_tp * 2 + (size_t)BIN_GAUSS * 16;
    SGS_CUDA_OK(set_max_smem(bin_count_kernel, smem_count));
    SGS_CUDA_OK(set_max_smem(bin_scatter_kernel, smem_scatter));
    SGS_CUDA_OK(launch_pdl(bin_count_kernel, lay.bin_ctas, BIN_THREADS, smem_count, stream, a));
    SGS_STAGE_OK(debug, stream);
    SGS_CUDA_OK(launch_pdl(bin_scan_kernel, lay.bin_tp / 64, SCAN_THREADS, 0, stream, a));
    SGS_STAGE_OK(debug, stream);
    SGS_CUDA_OK(launch_pdl(bin_scatter_kernel, lay.bin_ctas, BIN_THREADS, smem_scatter, stream, a));
    SGS_STAGE_OK(debug, stream);
    long long blocks = (L_cap + MASK_THREADS * MASK_PER - 1) / (MASK_THREADS * MASK_PER);
    if (blocks < 1) blocks = 1;
    SGS_CUDA_OK(launch_pdl(pair_masks_kernel, (unsigned)blocks, MASK_THREADS, 0, stream,
        (const unsigned long long*)a.keys, (const unsigned*)a.vals, (const int*)a.counters, L_cap,
        a.rec, lay.gx, a.masks));
    SGS_STAGE_OK(debug, stream);
    return 0;
}

}  // namespace sgs

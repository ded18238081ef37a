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
//   bin_scan   : per tile, exclusive prefix of the counts over the chunks -> base[c][t], total[t]
//   bin_scatter: tile starts = exclusive scan of total[] (a few thousand values, redone per CTA)
//                -> tile ranges, tiles bucketed by list length, pair count; then every pair of
//                the chunk goes to start[t] + base[c][t] + its rank inside the chunk.
//   pair_masks : one reach-mask byte per pair (which 8x4 pixel blocks of the tile it can touch).
// No radix pass over the pairs, no search for the ranges: the pair list is written once.
//
// Order inside a chunk.  Counting is order-free (shared-memory atomics).  Placement is not: a
// warp owns 128 consecutive Gaussians and takes them ONE PER STEP, lanes = the tiles of that
// Gaussian's rectangle (distinct, so plain read-modify-write of the warp's private per-tile
// counters); steps follow depth order, so ranks are stable by construction.  Everything a step
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
// shared memory, then walk the group again writing the prefixes. ----
__global__ void __launch_bounds__(SCAN_THREADS) bin_scan_kernel(BinArgs a) {
    __shared__ uint2 s_part[SCAN_THREADS / 32][32];
    pdl_sync();
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
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
}

// ---- scatter ----
__global__ void __launch_bounds__(BIN_THREADS) bin_scatter_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned char s_bin[];
    unsigned* s_base = reinterpret_cast<unsigned*>(s_bin);                                        // [tp]
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_bin + (size_t)a.tp * 4);          // [BIN_WARPS][tp]
    uint4* s_info = reinterpret_cast<uint4*>(s_bin + (size_t)a.tp * 4 + (size_t)BIN_WARPS * a.tp * 2);   // [BIN_WARPS][128]
    __shared__ unsigned s_wsum[BIN_WARPS];
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
    // ---- tile starts: exclusive scan of total[] (every CTA redoes it: tiles <= 8192 values) ----
    const int per = a.tp / BIN_THREADS;                          // consecutive tiles per thread (tp % 64 == 0, BIN_THREADS <= 64 ... see launch)
    const int t0 = tid * per;
    unsigned mine = 0;
    for (int k = 0; k < per; k += 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.total + t0 + k));
        mine += v.x + v.y + v.z + v.w;
    }
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += x;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();                                             // also: the counters are zero
    unsigned off = 0, grand = 0;
#pragma unroll
    for (int w = 0; w < BIN_WARPS; w++) {
        const unsigned v = s_wsum[w];
        if (w < warp) off += v;
        grand += v;
    }
    {
        unsigned start = off + incl - mine;
        const unsigned* brow = a.base + (size_t)chunk * a.tp;
        const unsigned cap = (unsigned)min(a.L_cap, (long long)0xffffffffll);
        for (int k = 0; k < per; k++) {
            const int t = t0 + k;
            const unsigned n = __ldg(a.total + t);
            s_base[t] = start + __ldg(brow + t);
            // ranges, buckets by list length: tile t is filed by the CTA it is congruent to
            if (t < a.tiles && (t % (int)gridDim.x) == chunk) {
                const unsigned lo = min(start, cap), hi = min(start + n, cap);
                a.ranges[t] = hi > lo ? make_uint2(lo, hi) : make_uint2(0u, 0u);
                const unsigned bk = (unsigned)(32 - __clz(hi - lo)) % 32u;
                const unsigned slot = atomicAdd(&a.bucket_count[bk], 1u);
                a.bucket_list[(size_t)bk * a.tiles + slot] = (unsigned)t;
            }
            start += n;
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
    // ---- pass B: place, one Gaussian per step in depth order, lanes = its tiles ----
    unsigned short* const cnt = s_cnt + (size_t)warp * a.tp;
    uint4 nx = info[0];
#pragma unroll 2
    for (int q = 0; q < BIN_WARP_GAUSS; q++) {
        const uint4 f = nx;
        if (q + 1 < BIN_WARP_GAUSS) nx = info[q + 1];
        const unsigned w = (f.x >> 16) & 0xffu, n = w * (f.x >> 24);
        for (unsigned k = lane; k < n; k += 32) {
            const unsigned ty = __umulhi(k, f.y), tx = k - ty * w;
            const unsigned t = ((f.x >> 8 & 0xffu) + ty) * (unsigned)a.gx + (f.x & 0xffu) + tx;
            const unsigned c = cnt[t];
            cnt[t] = (unsigned short)(c + 1u);
            const unsigned long long p = (unsigned long long)s_base[t] + c;
            if (p < (unsigned long long)a.L_cap) {
                a.keys[p] = ((unsigned long long)t << 32) | f.w;
                a.vals[p] = f.z;
            }
        }
        __syncwarp();                                            // the next step may hit the same tiles
    }
}

// ---- reach masks: one byte per pair of the sorted list ----
constexpr int MASK_THREADS = 256;
constexpr int MASK_PER = 2;
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
        const long long i = i0 + k * MASK_THREADS;
        tile[k] = i < n ? (unsigned)(__ldg(keys + i) >> 32) : 0u;
        id[k] = i < n ? __ldg(point_list + i) : 0u;
    }
    float4 q0[MASK_PER], q1[MASK_PER], q3[MASK_PER];
#pragma unroll
    for (int k = 0; k < MASK_PER; k++) {
        const float4* p = rec + 4 * (size_t)id[k];
        q0[k] = __ldg(p); q1[k] = __ldg(p + 1); q3[k] = __ldg(p + 3);
    }
#pragma unroll
    for (int k = 0; k < MASK_PER; k++) {
        const long long i = i0 + k * MASK_THREADS;
        if (i >= n) break;
        const float tx = (float)((tile[k] % (unsigned)gx_tiles) * TILE), ty = (float)((tile[k] / (unsigned)gx_tiles) * TILE);
        masks[i] = (unsigned char)reach_mask(q0[k], q1[k], q3[k], tx, ty);
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
        reinterpret_cast<const float4*>(geom + lay.rec_off), lay.gx,
        reinterpret_cast<unsigned char*>(bin + lay.masks_off)));
    SGS_STAGE_OK(debug, stream);
    return 0;
}

}  // namespace sgs

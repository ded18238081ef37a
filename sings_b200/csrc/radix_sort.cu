// radix_sort.cu -- hand-written onesweep LSD radix sort of (u64 key, u32 value) pairs.
//
// Replaces cub::DeviceRadixSort::SortPairs as called by the reference's rasterizer
// ([upstream] rasterizer_impl.cu: keys = tile<<32 | depth bits, values = Gaussian ids, bits
// [0, 32+getHigherMsb(tiles)); SURVEY.md A.1/A.3, K6 of section 2.4).  A stable LSD radix sort
// has a unique answer, so the sorted arrays are bit-identical to CUB's.
//
// In the rasterizer the 64-bit sort is split (raster_geometry.cu): the four depth digits are
// sorted once per GAUSSIAN (32-bit keys, launch_depth_sort; a pass whose digit is the same for
// every visible Gaussian is skipped), pairs are emitted in that order, and only the tile-id
// digits are sorted per PAIR (launch_tile_sort).  The stand-alone entry point sorts all digits.
//
// One kernel per 8-bit digit ("onesweep"): the digit histograms of ALL passes are produced
// up front (fused into the geometry / emission kernels, or by histogram_kernel for the
// stand-alone entry point); each pass ranks its tile stably (warp match_any + per-warp
// counters), obtains the tile's global digit offsets by decoupled look-back over the
// preceding tiles, and scatters through shared memory so global stores are digit-run
// coalesced.  Tiles are taken in ticket order (forward progress of the look-back).
// The item count is read from device memory, so the launch needs no host round trip.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace sgs {

constexpr unsigned ST_AGG = 1u << 30;
constexpr unsigned ST_INCL = 2u << 30;
constexpr unsigned ST_FLAG = 3u << 30;
constexpr unsigned ST_VAL = ~ST_FLAG;


// One pass.  K = unsigned (per-Gaussian depth items) or unsigned long long (pairs).
// The items come from buffer `par` of the ping-pong pair and go to the other one; with
// `varbits` set (depth sort) the pass first decides from the varying key bits whether it runs
// at all and which buffer is current (see depth_sort_parity, common.cuh).
template <typename K>
struct SortPass {
    K* keys[2];
    unsigned* vals[2];
    const unsigned* hist;     // 256 bins of this pass
    unsigned* status;         // [tiles][256] look-back words of this pass (zeroed)
    int* ticket;              // zeroed
    const int* n_ptr;         // item count on the device (may be null -> n_cap)
    const unsigned* varbits;  // null: always run, source = buffer `par`
    long long n_cap;
    int shift;
    unsigned mask;            // (1 << bits of this digit) - 1; < 255 only in a partial last pass
    int par;                  // source buffer (without varbits) / pass index (with varbits)
};

// exclusive scan of two values per thread over a 256-thread block (one barrier)
__device__ __forceinline__ uint2 block_excl_scan2(uint2 v, uint2* s_tmp, int lane, int warp) {
    uint2 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned nx = __shfl_up_sync(0xffffffffu, incl.x, d);
        unsigned ny = __shfl_up_sync(0xffffffffu, incl.y, d);
        if (lane >= d) { incl.x += nx; incl.y += ny; }
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    uint2 off = make_uint2(0u, 0u);
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++)
        if (w < warp) { off.x += s_tmp[w].x; off.y += s_tmp[w].y; }
    return make_uint2(off.x + incl.x - v.x, off.y + incl.y - v.y);
}

// look-back batch: predecessor words in flight per thread (A/B knob, tools/sweep.sh)
#ifndef SGS_SORT_EARLY_REORDER    // shared-memory reorder before (1) or after (0) the look-back
#define SGS_SORT_EARLY_REORDER 1
#endif
#ifndef SGS_LOOKBACK_N
#define SGS_LOOKBACK_N 4
#endif
#ifndef SGS_LOOKBACK_L
#define SGS_LOOKBACK_L 4
#endif

// resident CTAs per SM the pair-sort kernel is compiled for (register cap).  Measured (sweep
// v6b): 3 per SM -- a whole pass of 391 CTAs resident at once, no tickets -- is 9 us SLOWER per
// frame than 2 per SM with a second wave in ticket order.
#ifndef SGS_SORT_MINB_L
#define SGS_SORT_MINB_L 2
#endif

// Lanes of the warp holding the same 8-bit digit.  match.any's latency grows with the number of
// distinct values in the warp, and a warp of radix digits holds ~28 of them: eight ballots (one
// per bit, independent of each other) give the same mask in a few dozen cycles.
#ifndef SGS_SORT_MATCH_BALLOT
#define SGS_SORT_MATCH_BALLOT 1
#endif
__device__ __forceinline__ unsigned match_digit(unsigned d) {
#if SGS_SORT_MATCH_BALLOT
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
    }
    return peers;
#else
    return __match_any_sync(0xffffffffu, d);
#endif
}

template <typename K, int SORT_ITEMS, int LOOKBACK_BATCH>
__global__ void __launch_bounds__(SORT_THREADS, sizeof(K) == 8 ? SGS_SORT_MINB_L : 4) onesweep_pass_kernel(SortPass<K> a) {
    constexpr int SORT_TILE = SORT_ITEMS * SORT_THREADS;
    __shared__ unsigned s_cnt[SORT_THREADS / 32][RADIX];
    __shared__ unsigned s_bexcl[RADIX];
    __shared__ unsigned s_gbase[RADIX];
    extern __shared__ __align__(16) unsigned char s_dyn[];     // SORT_TILE keys, then SORT_TILE values
    K* const s_keys = reinterpret_cast<K*>(s_dyn);
    unsigned* const s_vals = reinterpret_cast<unsigned*>(s_dyn + (size_t)SORT_TILE * sizeof(K));
    __shared__ uint2 s_tmp2[SORT_THREADS / 32];
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) s_cnt[w][tid] = 0;
    pdl_sync();
    int src = a.par;
    if (a.varbits) {
        if (!depth_pass_runs(a.varbits[0] & a.varbits[1], a.par)) return;      // grid-uniform
        src = depth_sort_parity(a.varbits, a.par);
    }
    const K* __restrict__ keys_in = src ? a.keys[1] : a.keys[0];
    const unsigned* __restrict__ vals_in = src ? a.vals[1] : a.vals[0];
    K* __restrict__ keys_out = src ? a.keys[0] : a.keys[1];
    unsigned* __restrict__ vals_out = src ? a.vals[0] : a.vals[1];
    long long n = a.n_ptr ? (long long)*a.n_ptr : a.n_cap;
    if (n > a.n_cap) n = a.n_cap;
    if (a.ticket && tid == 0) s_tile = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int tile = a.ticket ? s_tile : (int)blockIdx.x;
    const long long start = (long long)tile * SORT_TILE;
    if (start >= n) return;
    const int n_valid = (int)min((long long)SORT_TILE, n - start);

    // ---- load, warp-striped: warp w owns items [w*256, w*256+256), item i of lane l = i*32+l ----
    K key[SORT_ITEMS];
    unsigned val[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int local = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        bool ok = local < n_valid;
        key[i] = ok ? keys_in[start + local] : (K)~(K)0;
        val[i] = ok ? vals_in[start + local] : 0u;
    }
    // ---- stable rank inside the warp's segment ----
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        unsigned d = (unsigned)(key[i] >> a.shift) & a.mask;
        unsigned peers = match_digit(d);
        int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (lane == leader) {
            old = s_cnt[warp][d];
            s_cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();
    // ---- per digit (thread d): exclusive scan over warps, tile count ----
    unsigned count = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
        unsigned t = s_cnt[w][tid];
        s_cnt[w][tid] = count;
        count += t;
    }
    // The tile's digit counts are published first; everything that needs only tile-local
    // information (the shared-memory reorder) is done while the other tiles publish theirs.
    unsigned* row = a.status + (size_t)tile * RADIX;
    st_relaxed_u32(&row[tid], (tile == 0 ? ST_INCL : ST_AGG) | count);
    // ---- digit bases: global (from the up-front histogram) and inside the tile ----
    const uint2 ex = block_excl_scan2(make_uint2(a.hist[tid], count), s_tmp2, lane, warp);
    s_bexcl[tid] = ex.y;
    __syncthreads();
#if SGS_SORT_EARLY_REORDER
    // ---- reorder through shared memory into tile-sorted order (keys/values leave registers) ----
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        unsigned d = (unsigned)(key[i] >> a.shift) & a.mask;
        unsigned pos = s_bexcl[d] + s_cnt[warp][d] + rank[i];
        s_keys[pos] = key[i];
        s_vals[pos] = val[i];
    }
#endif
    // ---- decoupled look-back over preceding tiles, one digit per thread ----
    // The predecessors' words are fetched LOOKBACK_BATCH at a time (independent loads in
    // flight together) and consumed in order: when every tile of a pass starts at once, tile
    // k finds its predecessors still at AGGREGATE, and a walk of one dependent L2 round trip
    // per predecessor would be the whole pass time.  With batches of B the INCLUSIVE front
    // moves B tiles per round trip while the walk comes B tiles per round trip towards it.
    unsigned prev = 0;
    if (tile > 0) {
        bool found = false;
        for (int j = tile - 1; !found; j -= LOOKBACK_BATCH) {
            unsigned s[LOOKBACK_BATCH];
#pragma unroll
            for (int k = 0; k < LOOKBACK_BATCH; k++)
                s[k] = j - k >= 0 ? ld_relaxed_u32(a.status + (size_t)(j - k) * RADIX + tid) : ST_INCL;
#pragma unroll
            for (int k = 0; k < LOOKBACK_BATCH; k++) {
                if (found) continue;
                unsigned v = s[k];
                while ((v & ST_FLAG) == 0) v = ld_relaxed_u32(a.status + (size_t)(j - k) * RADIX + tid);
                prev += v & ST_VAL;
                found = (v & ST_FLAG) == ST_INCL;
            }
        }
        st_relaxed_u32(&row[tid], ST_INCL | (prev + count));
    }
    s_gbase[tid] = ex.x + prev;
#if !SGS_SORT_EARLY_REORDER
    // ---- reorder through shared memory into tile-sorted order (keys/values leave registers) ----
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        unsigned d = (unsigned)(key[i] >> a.shift) & a.mask;
        unsigned pos = s_bexcl[d] + s_cnt[warp][d] + rank[i];
        s_keys[pos] = key[i];
        s_vals[pos] = val[i];
    }
#endif
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        int p = tid + k * SORT_THREADS;
        if (p < n_valid) {
            K kk = s_keys[p];
            unsigned d = (unsigned)(kk >> a.shift) & a.mask;
            size_t dst = (size_t)s_gbase[d] + (unsigned)(p - (int)s_bexcl[d]);
            const unsigned v = s_vals[p];
            keys_out[dst] = kk;
            vals_out[dst] = v;
        }
    }
}

// digit histograms of all passes in one read of the keys (stand-alone entry point only)
__global__ void __launch_bounds__(256) histogram_kernel(const unsigned long long* keys, long long n,
                                                        int passes, int end_bit, unsigned* hist) {
    __shared__ unsigned s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        unsigned long long k = keys[i];
        for (int p = 0; p < passes; p++) {
            int bits = min(RADIX_BITS, end_bit - p * RADIX_BITS);
            atomicAdd(&s_hist[p * RADIX + (unsigned)((k >> (p * RADIX_BITS)) & ((1u << bits) - 1u))], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// Tiles may be taken in blockIdx order instead of ticket order (one contended atomic and one
// L2 round trip less per CTA) when every CTA of the pass is resident at once: a CTA spinning
// on a predecessor's look-back word can then never keep that predecessor from running.
template <typename Kern>
static bool all_resident(Kern k, int blocks, size_t smem) {
    return (long long)blocks <= resident_ctas((const void*)k, SORT_THREADS, smem);
}

// passes [p0, p1) of the 64-bit pair sort; pass p sorts key bits [8p, 8p+8) below end_bit
static int run_passes(unsigned long long* k0, unsigned* v0, unsigned long long* k1, unsigned* v1,
                      const unsigned* hist, unsigned* status, int* tickets, const int* n_ptr,
                      long long n_cap, int p0, int p1, int end_bit, int blocks, cudaStream_t stream,
                      int debug) {
    for (int p = p0; p < p1; p++) {
        SortPass<unsigned long long> a;
        a.keys[0] = k0; a.keys[1] = k1;
        a.vals[0] = v0; a.vals[1] = v1;
        a.hist = hist + (size_t)p * RADIX;
        a.status = status + (size_t)(p - p0) * blocks * RADIX;
        a.ticket = tickets + p;
        a.n_ptr = n_ptr;
        a.varbits = nullptr;
        a.n_cap = n_cap;
        a.shift = p * RADIX_BITS;
        a.mask = (1u << min(RADIX_BITS, end_bit - p * RADIX_BITS)) - 1u;
        a.par = (p - p0) & 1;
        auto k = onesweep_pass_kernel<unsigned long long, SORT_ITEMS_L, SGS_LOOKBACK_L>;
        const size_t smem = (size_t)SORT_TILE_L * 12;
        SGS_CUDA_OK(set_max_smem(k, smem));
        if (all_resident(k, blocks, smem)) a.ticket = nullptr;
        SGS_CUDA_OK(launch_pdl(k, blocks, SORT_THREADS, smem, stream, a));
        SGS_STAGE_OK(debug, stream);
    }
    return 0;
}

int launch_depth_sort(int P, const RasterLayout& lay, char* bin, cudaStream_t stream, int debug) {
    if (P <= 0) return 0;
    int* counters = reinterpret_cast<int*>(bin + lay.cnt_off);
    for (int p = 0; p < DEPTH_PASSES; p++) {
        SortPass<unsigned> a;
        a.keys[0] = reinterpret_cast<unsigned*>(bin + lay.nkeys0_off);
        a.keys[1] = reinterpret_cast<unsigned*>(bin + lay.nkeys1_off);
        a.vals[0] = reinterpret_cast<unsigned*>(bin + lay.nvals0_off);
        a.vals[1] = reinterpret_cast<unsigned*>(bin + lay.nvals1_off);
        a.hist = reinterpret_cast<const unsigned*>(bin + lay.hist_off) + (size_t)p * RADIX;
        a.status = reinterpret_cast<unsigned*>(bin + lay.nstat_off) + (size_t)p * lay.nsort_blocks * RADIX;
        a.ticket = counters + CNT_SORT_TICKET0 + p;
        a.n_ptr = nullptr;
        a.varbits = reinterpret_cast<const unsigned*>(counters + CNT_VARBITS);
        a.n_cap = P;
        a.shift = p * RADIX_BITS;
        a.mask = RADIX - 1;
        a.par = p;
        auto k = onesweep_pass_kernel<unsigned, SORT_ITEMS_N, SGS_LOOKBACK_N>;
        const size_t smem = (size_t)SORT_TILE_N * 8;
        if (all_resident(k, lay.nsort_blocks, smem)) a.ticket = nullptr;
        SGS_CUDA_OK(launch_pdl(k, lay.nsort_blocks, SORT_THREADS, smem, stream, a));
        SGS_STAGE_OK(debug, stream);
    }
    return 0;
}

int launch_tile_sort(const RasterLayout& lay, long long L_cap, const char* geom, char* bin,
                     cudaStream_t stream, int debug) {
    if (L_cap >= (1ll << 30)) return SGS_ERR_CAPACITY;
    int* counters = reinterpret_cast<int*>(bin + lay.cnt_off);
    return run_passes(reinterpret_cast<unsigned long long*>(bin + lay.keys0_off),
                      reinterpret_cast<unsigned*>(bin + lay.vals0_off),
                      reinterpret_cast<unsigned long long*>(bin + lay.keys1_off),
                      reinterpret_cast<unsigned*>(bin + lay.vals1_off),
                      reinterpret_cast<const unsigned*>(bin + lay.hist_off),
                      reinterpret_cast<unsigned*>(bin + lay.sortstat_off),
                      counters + CNT_SORT_TICKET0, counters + CNT_NUM_RENDERED, L_cap, DEPTH_PASSES,
                      lay.passes, lay.end_bit, lay.sort_blocks, stream, debug);
}

size_t sort_scratch_bytes(long long n) {
    size_t blocks = (size_t)((n + SORT_TILE_L - 1) / SORT_TILE_L);
    if (blocks < 1) blocks = 1;
    return align_up(CNT_SLOTS * 4, 256) + align_up((size_t)MAX_PASSES * RADIX * 4, 256) +
           align_up((size_t)MAX_PASSES * blocks * RADIX * 4, 256);
}

int launch_sort_pairs_u64(unsigned long long* keys, unsigned* vals, unsigned long long* keys_tmp,
                          unsigned* vals_tmp, char* scratch, size_t scratch_bytes, long long n,
                          int end_bit, int* result_in_tmp, cudaStream_t stream) {
    if (n < 0 || end_bit < 1 || end_bit > 64 || n >= (1ll << 30)) return SGS_ERR_BAD_ARG;
    if (scratch_bytes < sort_scratch_bytes(n)) return SGS_ERR_CAPACITY;
    const int passes = (end_bit + RADIX_BITS - 1) / RADIX_BITS;
    const int blocks = (int)((n + SORT_TILE_L - 1) / SORT_TILE_L);
    if (result_in_tmp) *result_in_tmp = passes & 1;
    if (n == 0) return 0;
    SGS_CUDA_OK(cudaMemsetAsync(scratch, 0, sort_scratch_bytes(n), stream));
    int* tickets = reinterpret_cast<int*>(scratch);
    unsigned* hist = reinterpret_cast<unsigned*>(scratch + align_up(CNT_SLOTS * 4, 256));
    unsigned* status = reinterpret_cast<unsigned*>(scratch + align_up(CNT_SLOTS * 4, 256) +
                                                   align_up((size_t)MAX_PASSES * RADIX * 4, 256));
    int hb = (int)min((long long)148 * 8, (n + 255) / 256);
    histogram_kernel<<<hb, 256, 0, stream>>>(keys, n, passes, end_bit, hist);
    SGS_LAUNCH_OK();
    return run_passes(keys, vals, keys_tmp, vals_tmp, hist, status, tickets, nullptr, n, 0, passes,
                      end_bit, blocks, stream, 0);
}

}  // namespace sgs

// common.cuh -- shared device helpers for the sm_100a kernels of sings_b200.
//
// Built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
// -fmad=false is part of the numeric contract: the only fused multiply-adds are
// the explicit fmaf()/__fmaf_rn() calls, so the forward pass matches the C oracle bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgs {

constexpr int TILE = 16;            // 16x16 pixel tiles ([upstream] config.h BLOCK_X/BLOCK_Y)
constexpr int TILE_PIX = TILE * TILE;

// ---- error plumbing: entry points return 0, a negative argument error, or a cudaError_t ----
#define SGS_CUDA_OK(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) return (int)_e;              \
    } while (0)

#define SGS_LAUNCH_OK()                                     \
    do {                                                    \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return (int)_e;              \
    } while (0)

// debug mode: synchronise and check after each stage ([upstream] CHECK_CUDA(debug))
#define SGS_STAGE_OK(debug, stream)                         \
    do {                                                    \
        SGS_LAUNCH_OK();                                    \
        if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream)); \
    } while (0)

#ifndef SGS_ERR_BAD_ARG          // same values as include/sings_b200.h
#define SGS_ERR_BAD_ARG -1
#define SGS_ERR_BAD_SH_DEGREE -2
#define SGS_ERR_BAD_JOINTS -3
#define SGS_ERR_MISALIGNED -4
#define SGS_ERR_CAPACITY -5
#endif

// ---- counters block at the head of the zeroed scratch region (int32 slots) ----
enum : int {
    CNT_NUM_RENDERED = 0,   // L = sum of tiles touched (written by the geometry kernel)
    CNT_OVERFLOW = 1,       // set when L exceeded the binning capacity
    CNT_SCAN_TICKET = 2,    // block ticket of the geometry kernel
    CNT_VISIBLE = 3,        // number of Gaussians with radii > 0 (statistics)
    CNT_RANGES_DONE = 4,    // CTAs of the tile-range kernel that have finished
    CNT_VARBITS = 5,        // [5] = OR of the visible depth keys, [6] = OR of their complements
    CNT_SORT_TICKET0 = 8,   // + pass: block ticket of each radix pass (8 slots)
    CNT_BLEND_TICKET = 16,  // item ticket of the forward blend (persistent warps)
    CNT_SLOTS = 32,
};

// ---- programmatic dependent launch (PDL, sm_90+) ----
// Every kernel of the frame is launched with the programmatic-stream-serialization attribute
// and begins (after set-up that touches no global memory) with pdl_wait() followed by
// pdl_launch_dependents(): the next kernel's launch, CTA scheduling and prologue overlap this
// kernel's execution, and pdl_wait() returns only when the whole preceding grid has completed
// and its writes are visible.  EVERY kernel waits BEFORE it releases its dependents, so when
// kernel n+1 starts, kernel n-1 is complete: completion is transitive along the chain, and a
// kernel may read, ahead of its own wait, data that is at least two of OUR kernels old.
// The frame is ~20 kernels of 5-150 us: launch latency is a measurable share of it.
// SGS_NO_PDL=1 in the environment launches them the ordinary way (A/B, debugging).
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {       // the default kernel prologue
    pdl_wait();
    pdl_launch_dependents();
}

bool pdl_enabled();      // api.cu
// cudaFuncAttributeMaxDynamicSharedMemorySize, raised once per (kernel, device) under a mutex
// instead of on every launch (api.cu)
cudaError_t ensure_max_smem(const void* func, size_t bytes);
template <typename Kern>
inline cudaError_t set_max_smem(Kern k, size_t bytes) { return ensure_max_smem((const void*)k, bytes); }
// resident CTAs of a kernel on the current device (SMs x occupancy), cached under the same mutex
long long resident_ctas(const void* func, int threads, size_t smem);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Reach mask of a (tile, Gaussian) pair: bit w says whether the Gaussian's alpha >= 1/255
// footprint can touch 8x4 pixel block w of the tile (w = blend warp index): reaches_block for
// the 2 x 4 blocks, sharing the per-column / per-row terms.  Computed ONCE per frame and pair, by
// the emission kernel (one pair per lane, the Gaussian's record staged in shared memory), and
// packed into the top byte of the pair's 32-bit value -- so it rides through the tile-id radix
// passes for free, and the forward and the backward blend gather the 64-byte record only for
// the ~10 % of (warp, pair) combinations that can contribute.
__device__ __forceinline__ unsigned reach_mask(const float4 q0, const float4 q1, const float4 q3,
                                               float tx, float ty) {
    const float a = -2.0f * q0.z, b2 = -2.0f * q0.w, c = -2.0f * q1.x;
    float X0[2], X1[2], cx[2], acx2[2], bcx[2], ycx[2];
    float Y0[4], Y1[4], cy[4], ccy2[4], bcy[4], xcy[4];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        X0[k] = tx + (float)(8 * k) - q0.x; X1[k] = X0[k] + 7.0f;
        cx[k] = fminf(fmaxf(0.0f, X0[k]), X1[k]);
        acx2[k] = a * cx[k] * cx[k]; bcx[k] = b2 * cx[k]; ycx[k] = q3.y * cx[k];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        Y0[r] = ty + (float)(4 * r) - q0.y; Y1[r] = Y0[r] + 3.0f;
        cy[r] = fminf(fmaxf(0.0f, Y0[r]), Y1[r]);
        ccy2[r] = c * cy[r] * cy[r]; bcy[r] = b2 * cy[r]; xcy[r] = q3.z * cy[r];
    }
    unsigned m = 0;
#pragma unroll
    for (int w = 0; w < TILE_PIX / 32; w++) {
        const int k = w & 1, r = w >> 1;
        const float dy1 = fminf(fmaxf(ycx[k], Y0[r]), Y1[r]);       // minimiser on the edge x = cx
        const float dx2 = fminf(fmaxf(xcy[r], X0[k]), X1[k]);       // minimiser on the edge y = cy
        const float qa = fmaf(dy1, fmaf(c, dy1, bcx[k]), acx2[k]);
        const float qb = fmaf(dx2, fmaf(a, dx2, bcy[r]), ccy2[r]);
        if (fminf(qa, qb) <= q3.x) m |= 1u << w;
    }
    return m;
}


// ---- scoped loads/stores for the decoupled look-back status words ----
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- streaming (read-once) 128-bit global load, no L1 allocation ----
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// 128-bit read-only load that stays where it is written: the compiler may neither sink it to its
// first use nor merge it with a later one (software prefetch / memory-level parallelism)
__device__ __forceinline__ float4 ldg_f4_pinned(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned ldg_u32_pinned(const unsigned* p) {
    unsigned r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned long long ldg_u64_pinned(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// ---- cp.async (LDGSTS): 16-byte global -> shared, L2 only ----
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
// 16 bytes through L1 (a record gathered by one warp of a tile is wanted by its neighbours)
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
// 4 bytes, or 4 zero bytes when !valid (nothing is read from gmem_src then)
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem_src), "r"(valid ? 4 : 0) : "memory");
}
// all but the N most recently committed groups of this thread have landed
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA 1-D bulk copy (cp.async.bulk) + mbarrier: contiguous global chunk -> shared ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(s), "r"(parity)
        : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                             unsigned long long* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(gmem_src), "r"(bytes), "r"(b)
        : "memory");
}

// ---- vector reduction to global memory (sm_90+): one L2 atomic for four floats ----
__device__ __forceinline__ void red_add_f4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Depth-sort pass skipping.  varbits[0] & varbits[1] = the key bits that differ among the
// visible Gaussians; a pass whose 8-bit digit has none of them is the identity on the visible
// items and is not run.  Which of the two ping-pong buffers holds the items before pass
// `pass` (= after all passes when pass == DEPTH_PASSES) follows from the same two words.
__device__ __forceinline__ bool depth_pass_runs(unsigned diff, int pass) {
    return ((diff >> (8 * pass)) & 0xffu) != 0u;
}
__device__ __forceinline__ int depth_sort_parity(const unsigned* __restrict__ varbits, int pass) {
    const unsigned diff = varbits[0] & varbits[1];
    int par = 0;
    for (int q = 0; q < pass; q++) par ^= (int)depth_pass_runs(diff, q);
    return par;
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ unsigned lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// exp(x) for x <= 0, bit-identical to oracle/c/raster_oracle.c expneg (numeric contract).
// FMA/ALU pipes only (no MUFU, no F2I/FRND): round-to-nearest by the 1.5*2^23 magic add,
// Cody-Waite reduction, degree-5 polynomial in Estrin form (dependency depth 3), exponent
// spliced in with an integer add.  Max relative error 2.0e-7 on [-80, 0]; x < -80 is
// clamped (the result, 1.8e-35, is far below every threshold that consumes it).
__device__ __forceinline__ float expneg(float x) {
    x = fmaxf(x, -80.0f);
    const float r = __fmaf_rn(x, 1.44269504088896341f, 12582912.0f);
    const float n = __fadd_rn(r, -12582912.0f);
    float g = __fmaf_rn(n, -0.693359375f, x);
    g = __fmaf_rn(n, 2.12194440e-4f, g);
    const float g2 = __fmul_rn(g, g);
    const float a = __fmaf_rn(9.9999970198e-01f, g, 1.0f);
    const float b = __fmaf_rn(1.6667643189e-01f, g, 4.9999141693e-01f);
    const float c = __fmaf_rn(8.2901455462e-03f, g, 4.1898854077e-02f);
    const float p = __fmaf_rn(__fmaf_rn(c, g2, b), g2, a);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// ---- packed binary32 pairs (sm_100: FFMA2 / FMUL2 / FADD2, one issue slot for two IEEE rn
// operations).  The blend kernels evaluate two list entries per instruction with them; the
// results are those of the scalar sequence, lane by lane.  NOTE: ptxas contracts a
// mul.rn.f32x2 whose only consumer is an add.rn.f32x2 into one FFMA2 even under -fmad=false
// (the scalar _rn forms are never contracted), so bit-exact code must not spell that
// pattern: every multiply here feeds a multiply, an fma multiplicand, or an fma ADDEND. ----
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\n"
        "mul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\n"
        "add.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// expneg of two arguments at once (each lane of the pair: exactly the scalar sequence above).
// r2 is returned too: the caller splices the exponents in with integer adds.
__device__ __forceinline__ float2 expneg2(float2 x) {
    x.x = fmaxf(x.x, -80.0f); x.y = fmaxf(x.y, -80.0f);
    const float2 r = ffma2(x, splat2(1.44269504088896341f), splat2(12582912.0f));
    const float2 n = fadd2(r, splat2(-12582912.0f));
    float2 g = ffma2(n, splat2(-0.693359375f), x);
    g = ffma2(n, splat2(2.12194440e-4f), g);
    const float2 g2 = fmul2(g, g);
    const float2 a = ffma2(splat2(9.9999970198e-01f), g, splat2(1.0f));
    const float2 b = ffma2(splat2(1.6667643189e-01f), g, splat2(4.9999141693e-01f));
    const float2 c = ffma2(splat2(8.2901455462e-03f), g, splat2(4.1898854077e-02f));
    const float2 p = ffma2(ffma2(c, g2, b), g2, a);
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23)));
}

// column-major 4x4 times (x,y,z,1), row r  ([upstream] auxiliary.h transformPoint4x3/4x4)
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
    float t = __fmul_rn(m[r], x);
    t = __fmaf_rn(m[4 + r], y, t);
    t = __fmaf_rn(m[8 + r], z, t);
    return __fadd_rn(t, m[12 + r]);
}

}  // namespace sgs

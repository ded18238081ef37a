// emu.h -- a minimal SIMT emulation for CPU-side tests of simple CUDA kernels (TEST INFRASTRUCTURE;
// never part of the product: the library has no CPU path).  A kernel source is compiled by g++ with
// this header standing in for the CUDA runtime: every CUDA thread of a block is a host thread,
// __syncthreads() is a barrier over the block, warp shuffles go through an exchange buffer (all threads
// of a WARP must execute the same shuffles -- a warp may retire as a whole before them),
// blocks run one after the other (so `__shared__` can be a plain static).  tests/cuda_emu/__init__.py
// rewrites `kernel<<<grid, block, smem, stream>>>(args)` into EMU_LAUNCH(...).
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

struct uint3_ { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct float3 { float x, y, z; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { return uint3{x, y, z}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
// round-to-nearest single operations (the build passes -ffp-contract=off) and the conversions used by the kernels
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fdividef(float a, float b) { return a / b; }       // (approximate on the device)
inline float __logf(float a) { return logf(a); }                  // (approximate on the device)
inline int __float2int_rz(float x) { return (int)x; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T max(T a, T b) { return a < b ? b : a; }
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __cvta_generic_to_shared(p) ((size_t)(p))
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
struct Block {
    pthread_barrier_t bar;                       // __syncthreads
    std::vector<pthread_barrier_t> warp_bar;     // warp-synchronous primitives: one barrier per warp
    std::vector<unsigned long long> xchg;
    unsigned threads;
    char* dyn_smem;                              // `extern __shared__` of this launch (128-byte aligned)
};
inline thread_local Block* blk = nullptr;
inline std::mutex atomic_mutex;
// emulated mbarrier word: bit 63 = parity of the phase in progress, low bits = bytes still expected
inline void mbar_complete_if_done(unsigned long long* bar) {
    const unsigned long long v = __atomic_load_n(bar, __ATOMIC_SEQ_CST);
    if ((v & 0x7fffffffffffffffull) == 0) __atomic_store_n(bar, (v ^ 0x8000000000000000ull) & 0x8000000000000000ull, __ATOMIC_SEQ_CST);
}
}  // namespace emu

inline thread_local uint3_ threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { pthread_barrier_wait(&emu::blk->bar); }

namespace emu {
inline unsigned linear_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
inline void warp_sync() { pthread_barrier_wait(&blk->warp_bar[linear_tid() >> 5]); }
// every lane of the warp publishes a value, then reads the one of lane `src` (all lanes of the warp that are
// still alive must take part: the kernels emulated here keep their warps converged around shuffles)
template <typename T>
inline T exchange(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const unsigned tid = linear_tid();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    blk->xchg[tid] = raw;
    warp_sync();
    const unsigned src = (tid & ~31u) | (src_lane & 31u);
    raw = blk->xchg[src < blk->threads ? src : tid];
    warp_sync();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
}  // namespace emu

inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) { return emu::exchange(v, (unsigned)src_lane); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu::exchange(v, (emu::linear_tid() ^ (unsigned)lane_mask) & 31u); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const unsigned lane = emu::linear_tid() & 31u;
    return emu::exchange(v, lane + delta < 32u ? lane + delta : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned tid = emu::linear_tid(), base = tid & ~31u;
    emu::blk->xchg[tid] = pred ? 1ull : 0ull;
    emu::warp_sync();
    unsigned bits = 0;
    for (unsigned l = 0; l < 32 && base + l < emu::blk->threads; l++) bits |= (unsigned)(emu::blk->xchg[base + l] & 1ull) << l;
    emu::warp_sync();
    return bits;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const unsigned lane = emu::linear_tid() & 31u;
    return emu::exchange(v, lane >= delta ? lane - delta : lane);
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    const unsigned tid = emu::linear_tid(), base = tid & ~31u;
    emu::blk->xchg[tid] = v;
    emu::warp_sync();
    unsigned bits = 0;
    for (unsigned l = 0; l < 32 && base + l < emu::blk->threads; l++) bits |= (unsigned)((unsigned)emu::blk->xchg[base + l] == v) << l;
    emu::warp_sync();
    return bits;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    const unsigned tid = emu::linear_tid(), base = tid & ~31u;
    emu::blk->xchg[tid] = v;
    emu::warp_sync();
    unsigned r = 0;
    for (unsigned l = 0; l < 32 && base + l < emu::blk->threads; l++) r |= (unsigned)emu::blk->xchg[base + l];
    emu::warp_sync();
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) { const unsigned w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}

template <typename T>
inline T atomicAdd(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    *p = old + v;
    return old;
}

template <typename T>
inline T atomicMax(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    if (v > old) *p = v;
    return old;
}

template <typename T>
inline T atomicMin(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    if (v < old) *p = v;
    return old;
}
template <typename T>
inline T atomicOr(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    *p = old | v;
    return old;
}
template <typename T>
inline T atomicExch(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    *p = v;
    return old;
}

inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// ---- the slice of the CUDA runtime API the library's host code touches: everything is synchronous here ----
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeRelaxed = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 3 };
constexpr unsigned cudaEventRecordExternal = 1;
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };
constexpr cudaError_t cudaErrorNotSupported = 801;
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
template <typename T>
inline cudaError_t cudaMemcpyFromSymbol(void* d, const T& sym, size_t n) { memcpy(d, &sym, n); return cudaSuccess; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t, cudaStream_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = cudaStreamCaptureStatusNone; return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long = 0) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }

namespace emu {
template <typename F, typename... Args>
void launch_smem(F kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
    const unsigned nthreads = block.x * block.y * block.z;
    void* dyn = nullptr;
    if (posix_memalign(&dyn, 128, smem + 128) != 0) abort();
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block b;
                b.threads = nthreads;
                b.dyn_smem = static_cast<char*>(dyn);
                memset(dyn, 0xCD, smem + 128);       // shared memory starts undefined
                b.xchg.assign(nthreads, 0);
                pthread_barrier_init(&b.bar, nullptr, nthreads);
                b.warp_bar.resize((nthreads + 31) / 32);
                for (unsigned w = 0; w < b.warp_bar.size(); w++)
                    pthread_barrier_init(&b.warp_bar[w], nullptr, nthreads - 32 * w < 32 ? nthreads - 32 * w : 32);
                std::vector<std::thread> pool;
                pool.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; t++)
                    pool.emplace_back([=, &b]() {
                        blk = &b;
                        threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        blockIdx = {bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        kernel(args...);
                    });
                for (auto& th : pool) th.join();
                pthread_barrier_destroy(&b.bar);
                for (auto& wb : b.warp_bar) pthread_barrier_destroy(&wb);
            }
    free(dyn);
}
template <typename F, typename... Args>
void launch(F kernel, dim3 grid, dim3 block, Args... args) { launch_smem(kernel, grid, block, (size_t)0, args...); }
}  // namespace emu

template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args... args) {
    emu::launch_smem(kernel, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, static_cast<KArgs>(args)...);
    return cudaSuccess;
}

#define EMU_LAUNCH(kernel, grid, block, smem, ...) emu::launch_smem(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)

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
inline int __float2int_rz(float x) { return (int)x; }
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
struct Block {
    pthread_barrier_t bar;                       // __syncthreads
    std::vector<pthread_barrier_t> warp_bar;     // warp-synchronous primitives: one barrier per warp
    std::vector<unsigned long long> xchg;
    unsigned threads;
};
inline thread_local Block* blk = nullptr;
inline std::mutex atomic_mutex;
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
    unsigned bits = 0;
    for (unsigned l = 0; l < 32; l++) bits |= (emu::exchange<unsigned>(pred ? 1u : 0u, l) & 1u) << l;
    return bits;
}
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <typename T>
inline T __ldg(const T* p) { return *p; }

template <typename T>
inline T atomicAdd(T* p, T v) {
    std::lock_guard<std::mutex> g(emu::atomic_mutex);
    const T old = *p;
    *p = old + v;
    return old;
}

inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

namespace emu {
template <typename F, typename... Args>
void launch(F kernel, dim3 grid, dim3 block, Args... args) {
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block b;
                b.threads = nthreads;
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
}
}  // namespace emu

#define EMU_LAUNCH(kernel, grid, block, ...) emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)

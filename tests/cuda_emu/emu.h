// emu.h -- a minimal SIMT emulation for CPU-side tests of simple CUDA kernels (TEST INFRASTRUCTURE;
// never part of the product: the library has no CPU path).  A kernel source is compiled by g++ with
// this header standing in for the CUDA runtime: every CUDA thread of a block is a host thread,
// __syncthreads() is a barrier over the block, warp shuffles go through an exchange buffer (all threads
// of the block must execute the same shuffles -- true for the block reductions this is used for),
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
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

namespace emu {
struct Block {
    pthread_barrier_t bar;
    std::vector<unsigned long long> xchg;
    unsigned threads;
};
inline thread_local Block* blk = nullptr;
inline std::mutex atomic_mutex;
}  // namespace emu

inline thread_local uint3_ threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { pthread_barrier_wait(&emu::blk->bar); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const unsigned tid = threadIdx.x;
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::blk->xchg[tid] = raw;
    pthread_barrier_wait(&emu::blk->bar);
    const unsigned src = (tid & ~31u) | ((tid ^ (unsigned)lane_mask) & 31u);
    raw = emu::blk->xchg[src < emu::blk->threads ? src : tid];
    pthread_barrier_wait(&emu::blk->bar);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}

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
            }
}
}  // namespace emu

#define EMU_LAUNCH(kernel, grid, block, ...) emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__)

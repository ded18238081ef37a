"""Compile a simple CUDA kernel source for the HOST against tests/cuda_emu/emu.h (every CUDA thread a host
thread) so that `-m "not gpu"` tests can execute the kernel code itself, not a restatement of it.
TEST INFRASTRUCTURE: the product has no CPU path; nothing under sings_b200/ imports this."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

_COMMON = '''#pragma once
#include "emu.h"
#define SGS_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (int)_e; } while (0)
#define SGS_LAUNCH_OK() do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return (int)_e; } while (0)
#define SGS_ERR_BAD_ARG -1
#define SGS_ERR_BAD_SH_DEGREE -2
#define SGS_ERR_BAD_JOINTS -3
#define SGS_ERR_MISALIGNED -4
#define SGS_ERR_CAPACITY -5
namespace sgs {
// stand-ins for the helpers of the real common.cuh that are PTX there: programmatic dependent launch
// (ordering only), the streaming load (a cache hint)
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
inline void pdl_sync() {}
inline float4 ldg_stream_f4(const float4* p) { return *p; }
// packed-pair arithmetic (fma / mul / add .rn.f32x2 in the real header): two independent IEEE operations
inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 fadd2(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 splat2(float v) { return float2{v, v}; }
inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t, cudaStream_t, Args&&... args) {
    emu::launch(kernel, grid, block, static_cast<KArgs>(args)...);
    return cudaSuccess;
}
}  // namespace sgs
'''

_LAUNCH = re.compile(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;]*>)?)\s*<<<\s*([^;]*?)>>>\s*\(")


def _rewrite_launches(src: str) -> str:
    """kernel<T><<<grid, block, smem, stream>>>(args...)  ->  EMU_LAUNCH((kernel<T>), grid, block, args...)"""
    out, pos = [], 0
    for m in _LAUNCH.finditer(src):
        cfg = [c.strip() for c in m.group(2).split(",")]
        if len(cfg) < 2:
            raise ValueError(f"launch configuration not understood: {m.group(0)}")
        out.append(src[pos:m.start()])
        out.append(f"EMU_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, ")
        pos = m.end()
    out.append(src[pos:])
    return "".join(out)


def build(cu_path: str, exports: str) -> ctypes.CDLL:
    """g++-compile `cu_path` with its launches rewritten plus `exports` (extern "C" wrappers appended to the
    translation unit) into a shared object and load it."""
    src = _rewrite_launches(open(cu_path).read()) + "\n" + exports
    tag = hashlib.sha1((src + open(os.path.join(HERE, "emu.h")).read()).encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), f"sgs_cuda_emu_{tag}")
    so = os.path.join(d, "kernel_emu.so")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, "common.cuh"), "w").write(_COMMON)
        open(os.path.join(d, "kernels.h"), "w").write("#pragma once\n")
        cpp = os.path.join(d, "kernel_emu.cpp")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                               "-I", HERE, "-I", d, cpp, "-o", so + ".tmp"])
        os.replace(so + ".tmp", so)
    return ctypes.CDLL(so)

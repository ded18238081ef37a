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

# Emulation bodies for the helpers of sings_b200/csrc/common.cuh that are inline PTX there.  Everything else of
# that header (numeric helpers such as expneg / reach_mask / xform_row, enums, launch_pdl) is compiled as it is.
_EMU_BODIES = {
    "pdl_launch_dependents": "{}",
    "pdl_wait": "{}",
    "ld_relaxed_u64": "{ return __atomic_load_n(p, __ATOMIC_SEQ_CST); }",
    "st_relaxed_u64": "{ __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }",
    "ld_relaxed_u32": "{ return __atomic_load_n(p, __ATOMIC_SEQ_CST); }",
    "st_relaxed_u32": "{ __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }",
    "ldg_stream_f4": "{ return *p; }",
    "ldg_f4_pinned": "{ return *p; }",
    "ldg_u32_pinned": "{ return *p; }",
    "ldg_u64_pinned": "{ return *p; }",
    # cp.async: the copy happens at once (a legal execution: data only has to be there after the wait)
    "cp_async16": "{ memcpy(smem_dst, gmem_src, 16); }",
    "cp_async4": "{ memcpy(smem_dst, gmem_src, 4); }",
    "cp_async16_ca": "{ memcpy(smem_dst, gmem_src, 16); }",
    "cp_async4_zfill": "{ if (valid) memcpy(smem_dst, gmem_src, 4); else memset(smem_dst, 0, 4); }",
    "cp_async_wait_group": "{}",
    "cp_async_commit": "{}",
    "cp_async_wait_all": "{}",
    # mbarrier + TMA bulk load, for the one pattern the kernels use: thread 0 arms the barrier with the byte total
    # and issues the copies (synchronous here), every thread then waits for the phase.  Emulated word: bit 63 =
    # parity of the phase in progress, low bits = bytes still expected.
    "mbar_init": "{ (void)count; __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }",
    "mbar_fence_init": "{}",
    "fence_proxy_async": "{}",
    "mbar_arrive_expect_tx": "{ __atomic_fetch_add(bar, (unsigned long long)bytes, __ATOMIC_SEQ_CST); if (bytes == 0) emu::mbar_complete_if_done(bar); }",
    "mbar_wait": "{ while (((__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 63) & 1ull) == (unsigned long long)(parity & 1u)) std::this_thread::yield(); }",
    "tma_bulk_g2s": "{ memcpy(smem_dst, gmem_src, bytes); __atomic_fetch_sub(bar, (unsigned long long)bytes, __ATOMIC_SEQ_CST); emu::mbar_complete_if_done(bar); }",
    "red_add_f4": "{ std::lock_guard<std::mutex> g(emu::atomic_mutex); addr[0] += a; addr[1] += b; addr[2] += c; addr[3] += d; }",
    "lanemask_lt": "{ return (1u << (emu::linear_tid() & 31u)) - 1u; }",
    "lanemask_le": "{ const unsigned l = emu::linear_tid() & 31u; return ((1u << l) - 1u) | (1u << l); }",
    # packed pairs: two independent IEEE round-to-nearest operations
    "ffma2": "{ return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }",
    "fmul2": "{ return float2{a.x * b.x, a.y * b.y}; }",
    "fadd2": "{ return float2{a.x + b.x, a.y + b.y}; }",
}

_FUNC_HEAD = re.compile(r"(?:template\s*<[^>]*>\s*)?__device__\s+__forceinline__\s+[\w\s\*]+?\b(\w+)\s*\(([^)]*)\)\s*\{")


def transform_common(src: str) -> str:
    """sings_b200/csrc/common.cuh with every function whose body is inline PTX given its emulation body."""
    out, pos, seen = [], 0, set()
    for m in _FUNC_HEAD.finditer(src):
        if m.start() < pos:
            continue
        depth, k = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        body = src[m.end() - 1:k]
        if "asm" not in body:
            continue
        name = m.group(1)
        if name not in _EMU_BODIES:
            raise KeyError(f"common.cuh: no emulation body for the inline-PTX helper {name}()")
        seen.add(name)
        out.append(src[pos:m.end() - 1])
        out.append(_EMU_BODIES[name])
        pos = k
    out.append(src[pos:])
    res = "".join(out)
    if "asm" in re.sub(r"//[^\n]*", "", res):
        raise ValueError("common.cuh: an inline-PTX statement survived the transformation")
    return res


# host-side helpers that api.cu defines (kernel attribute caches); single-file builds get these instead
_HOST_STUBS = """
namespace sgs {
bool pdl_enabled() { return false; }
cudaError_t ensure_max_smem(const void*, size_t) { return cudaSuccess; }
long long resident_ctas(const void*, int, size_t) { return 296; }
}
"""

_LAUNCH = re.compile(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;]*>)?)\s*<<<\s*([^;]*?)>>>\s*\(")


def _rewrite_launches(src: str) -> str:
    """kernel<T><<<grid, block, smem, stream>>>(args...)  ->  EMU_LAUNCH((kernel<T>), grid, block, smem, args...)"""
    out, pos = [], 0
    for m in _LAUNCH.finditer(src):
        cfg, depth, cur = [], 0, ""
        for ch in m.group(2):                      # split at top-level commas only: <<<min(a, b), 256, 0, stream>>>
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                cfg.append(cur.strip()); cur = ""
            else:
                cur += ch
        cfg.append(cur.strip())
        if len(cfg) < 2:
            raise ValueError(f"launch configuration not understood: {m.group(0)}")
        out.append(src[pos:m.start()])
        out.append(f"EMU_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, {cfg[2] if len(cfg) > 2 else 0}, ")
        pos = m.end()
    out.append(src[pos:])
    return "".join(out)


_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")
CSRC = os.path.join(ROOT, "sings_b200", "csrc")
# SGS_EMU_SANITIZE=address|thread: build the emulated kernels with that sanitizer (run the tests with the matching
# runtime preloaded, tools/emu_sanitize.sh) -- out-of-bounds accesses / data races inside a block in the kernel
# source are then reported on the CPU
_SAN = os.environ.get("SGS_EMU_SANITIZE", "")
_GXX = ["g++", "-std=c++17", "-O1", "-fPIC", "-pthread", "-ffp-contract=off", "-w"] + \
    ([f"-fsanitize={_SAN}", "-g", "-fno-omit-frame-pointer"] if _SAN else [])


def _prepare(text: str, rewrites=()) -> str:
    for pat, rep in rewrites:
        text, n = re.subn(pat, rep, text)
        if n == 0:
            raise ValueError(f"rewrite {pat!r} matched nothing")
    text = _DYN_SMEM.sub(r"\1* \2 = reinterpret_cast<\1*>(emu::blk->dyn_smem);", text)
    return _rewrite_launches(text)


def _stage_headers(d: str) -> None:
    """The library's own headers beside the translation units: common.cuh transformed, the others as they are;
    <cuda_runtime.h> resolves to the emulation."""
    open(os.path.join(d, "cuda_runtime.h"), "w").write('#pragma once\n#include "emu.h"\n')
    for h in sorted(os.listdir(CSRC)):
        if h.endswith((".h", ".cuh")):
            text = open(os.path.join(CSRC, h)).read()
            text = transform_common(text) if h == "common.cuh" else text
            open(os.path.join(d, h), "w").write(_DYN_SMEM.sub(r"\1* \2 = reinterpret_cast<\1*>(emu::blk->dyn_smem);", text))
    inc = os.path.join(ROOT, "include", "sings_b200.h")
    os.makedirs(os.path.join(d, "include"), exist_ok=True)
    open(os.path.join(d, "include", "sings_b200.h"), "w").write(open(inc).read())


def _digest(*texts) -> str:
    h = hashlib.sha1()
    for t in texts:
        h.update(t.encode())
    for f in sorted(os.listdir(CSRC)) + ["emu.h"]:
        path = os.path.join(CSRC, f) if f != "emu.h" else os.path.join(HERE, f)
        if f.endswith((".h", ".cuh")):
            h.update(open(path).read().encode())
    h.update(open(os.path.abspath(__file__)).read().encode())
    h.update(_SAN.encode())
    return h.hexdigest()[:16]


# the three inline-PTX statements lbs.cu defines itself (bulk shared -> global store and its group bookkeeping)
LBS_REWRITES = [(r'asm volatile\("cp\.async\.bulk\.global\.shared::cta\.bulk_group[^;]*;"[^;]*;', "memcpy(gmem_dst, smem_src, bytes); (void)sa;"),
                (r'asm volatile\("cp\.async\.bulk\.commit_group;" ::: "memory"\);', "(void)0;"),
                (r'asm volatile\("cp\.async\.bulk\.wait_group\.read 0;" ::: "memory"\);', "(void)0;")]


_THREADS_OK = None


def require_threads(n: int = 300) -> None:
    """The emulation runs one host thread per CUDA thread of a block (up to 256 + the caller): skip the test
    (pytest.skip) where the environment cannot start that many, instead of aborting inside std::thread."""
    global _THREADS_OK
    if _THREADS_OK is None:
        import threading
        gate, started = threading.Event(), []
        try:
            for _ in range(n):
                t = threading.Thread(target=gate.wait)
                t.start()
                started.append(t)
            _THREADS_OK = True
        except RuntimeError:
            _THREADS_OK = False
        finally:
            gate.set()
            for t in started:
                t.join()
    if not _THREADS_OK:
        import pytest
        pytest.skip(f"cannot start {n} threads here (the SIMT emulation needs one per CUDA thread of a block)")


def build(cu_path: str, exports: str, rewrites=(), headers=()) -> ctypes.CDLL:
    """g++-compile ONE kernel source (launches rewritten, `extern "C"` wrappers `exports` appended to the
    translation unit) against the emulation into a shared object and load it.  `rewrites`: (regex, replacement)
    pairs applied to the source first (for the few inline-PTX statements a file defines itself)."""
    require_threads()
    src = _prepare(open(cu_path).read(), rewrites) + "\n" + _HOST_STUBS + "\n" + exports
    d = os.path.join(tempfile.gettempdir(), f"sgs_cuda_emu_{_digest(src)}")
    so = os.path.join(d, "kernel_emu.so")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        _stage_headers(d)
        cpp = os.path.join(d, "kernel_emu.cpp")
        open(cpp, "w").write(src)
        subprocess.check_call(_GXX + ["-shared", "-I", HERE, "-I", d, cpp, "-o", so + ".tmp"])
        os.replace(so + ".tmp", so)
    return ctypes.CDLL(so)


def build_library() -> ctypes.CDLL:
    """The WHOLE library -- every .cu of sings_b200/csrc incl. api.cu, i.e. the real C ABI of
    include/sings_b200.h -- compiled against the emulation: one object per source, linked into one shared
    object.  CUDA graphs are not available in it (capture returns an error); everything else runs."""
    require_threads()
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    special = {"lbs.cu": LBS_REWRITES, "api.cu": [(r'#include "\.\./\.\./include/sings_b200\.h"', '#include "include/sings_b200.h"')]}
    texts = {f: _prepare(open(os.path.join(CSRC, f)).read(), special.get(f, ())) for f in srcs}
    d = os.path.join(tempfile.gettempdir(), f"sgs_cuda_emu_lib_{_digest(*[texts[f] for f in srcs])}")
    so = os.path.join(d, "libsings_b200_emu.so")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        _stage_headers(d)
        procs, objs = [], []
        for f in srcs:
            cpp = os.path.join(d, f[:-3] + ".cpp")
            open(cpp, "w").write(texts[f])
            objs.append(cpp[:-4] + ".o")
            procs.append((f, subprocess.Popen(_GXX + ["-c", "-I", HERE, "-I", d, cpp, "-o", objs[-1]], stderr=subprocess.PIPE)))
        for f, pr in procs:
            err = pr.communicate()[1]
            if pr.returncode:
                raise RuntimeError(f"emulated build of {f} failed:\n" + err.decode()[-6000:])
        subprocess.check_call(["g++", "-shared", "-pthread"] + ([f"-fsanitize={_SAN}"] if _SAN else []) + objs + ["-o", so + ".tmp"])
        os.replace(so + ".tmp", so)
    return ctypes.CDLL(so)

"""CPU: the C-ABI library loads and exports every symbol include/sings_b200.h declares
(no compute calls without a GPU); host-side helpers that need no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sings_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgs_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sings_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/sings_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), "python binding table and header disagree"


def test_version_error_strings_and_sizes_need_no_gpu():
    from sings_b200 import _lib
    L = _lib.lib()
    assert L.sgs_version() >= 100
    assert b"bad argument" in L.sgs_error_string(-1)
    assert b"SH degree" in L.sgs_error_string(-2)
    from sings_b200.rasterizer import _sizes, layout_info
    g, b, i, a = _sizes(1000, 640, 480, 50_000)
    assert g >= 1000 * 48 and b >= 50_000 * 24 and i >= 640 * 480 * 8 and a >= 1000 * 48
    info = layout_info(1000, 1024, 1024, 50_000)
    assert info["tiles"] == 4096 and info["end_bit"] == 45 and info["passes"] == 6
    info = layout_info(10, 1920, 1080, 1000)
    assert info["tiles"] == 120 * 68 and info["end_bit"] == 45
    info = layout_info(10, 2048, 2048, 1000)
    assert info["end_bit"] == 47


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing routes through the oracle."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from sings_b200 import _lib, deform
    rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                       torch.zeros(3), False, False)
    with pytest.raises(_lib.SgsError):
        GaussianRasterizer(rs)(means3D=torch.zeros(4, 3), means2D=torch.zeros(4, 3), opacities=torch.ones(4, 1),
                               colors_precomp=torch.ones(4, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(_lib.SgsError):
        deform.deform_gaussians(torch.eye(4).expand(24, 4, 4), torch.zeros(4, 3), torch.ones(4, 24) / 24, None,
                                torch.ones(4, 3))
    for mod in ("sings_b200/rasterizer.py", "sings_b200/deform.py", "sings_b200/step.py", "sings_b200/dp.py",
                "sings_b200/_lib.py", "diff_gaussian_rasterization/__init__.py"):
        src = open(os.path.join(ROOT, mod)).read()
        assert "oracle" not in src, f"{mod} must not reference the oracle"


def test_settings_tuple_is_the_classic_12_field_api():
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    f = GaussianRasterizationSettings._fields
    assert f[:12] == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                      "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    # constructible with exactly the 12 keywords the reference passes (gs_renderer_single.py:69-82)
    GaussianRasterizationSettings(image_height=1, image_width=1, tanfovx=1.0, tanfovy=1.0, bg=None,
                                  scale_modifier=1.0, viewmatrix=None, projmatrix=None, sh_degree=0,
                                  campos=None, prefiltered=False, debug=False)


def test_flag_constants_match_header():
    """The bits of the rasterizer entry points' `debug` argument are the same in the header and in
    the Python binding."""
    import re
    from sings_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "sings_b200.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SGS_FLAG_(\w+)\s+(\d+)", hdr)}
    assert flags == {"SYNC_CHECK": _lib.FLAG_SYNC_CHECK, "PRECLEARED": _lib.FLAG_PRECLEARED,
                     "EARLY_PARAMS": _lib.FLAG_EARLY_PARAMS, "FORWARD_ONLY": _lib.FLAG_FORWARD_ONLY}


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/sings_b200.h compiles as C99 (and C++17) with no torch or CUDA headers, and a C program
    that takes the address of every declared entry point links against the shared library: the
    boundary a non-Python host (cgo / JNI / FFI) binds is exactly this header."""
    import re
    import shutil
    import subprocess
    from sings_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.join(os.path.dirname(__file__), "..")
    hdr = open(os.path.join(root, "include", "sings_b200.h")).read()
    names = sorted(set(re.findall(r"\b(sgs_[A-Za-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S))))
    assert set(names) >= set(_lib.EXPORTS)
    src = os.path.join(tmp_path, "bind.c")
    with open(src, "w") as f:
        f.write('#include "sings_b200.h"\n#include <stdio.h>\nint main(void) {\n  const void* fns[] = {\n')
        f.write("".join(f"    (const void*)&{n},\n" for n in names))
        f.write('  };\n  printf("%d %d\\n", (int)(sizeof(fns) / sizeof(fns[0])), sgs_version());\n  return 0;\n}\n')
    inc = os.path.join(root, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-Wno-pedantic", "-fsyntax-only", "-I", inc, src])
    if shutil.which("g++"):
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", inc, "-x", "c++", src])
    _lib.build()
    libdir = os.path.dirname(_lib.lib_path()) if hasattr(_lib, "lib_path") else os.path.join(root, "sings_b200", "lib")
    exe = os.path.join(tmp_path, "bind")
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, src, "-o", exe, "-L", libdir, "-lsings_b200", f"-Wl,-rpath,{os.path.abspath(libdir)}"])
    out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == len(names) and int(out[1]) >= 200

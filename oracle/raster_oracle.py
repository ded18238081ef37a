"""ctypes/numpy front end of oracle/c/raster_oracle.c (CPU restatement of the rasterizer).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by sings_b200/ or diff_gaussian_rasterization/.

Parity status: "parity unpinned" (see the header of raster_oracle.c): the restated algorithm
is the un-vendored, unpinned diff-gaussian-rasterization that the reference installs at
/root/reference/install_all.sh:22 and calls at
/root/reference/sings/rec/renderer/gs_renderer_single.py:69-95.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraster_oracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "c", "raster_oracle.c")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ro_expneg.restype = C.c_float
        _lib.ro_expneg.argtypes = [C.c_float]
        _lib.ro_scan.restype = C.c_int64
        _lib.ro_higher_msb.restype = C.c_int
        _lib.ro_max_threads.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray], typ):
    if a is None:
        return typ()
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(typ)


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def set_threads(n: int) -> None:
    lib().ro_set_threads(C.c_int(n))


def max_threads() -> int:
    return int(lib().ro_max_threads())


def expneg(x: float) -> float:
    return float(lib().ro_expneg(C.c_float(x)))


def higher_msb(n: int) -> int:
    return int(lib().ro_higher_msb(C.c_uint32(n)))


@dataclass
class Camera:
    """The camera part of GaussianRasterizationSettings (gs_renderer_single.py:69-82)."""

    W: int
    H: int
    tanfovx: float
    tanfovy: float
    view: np.ndarray   # 16 floats, column-major == torch (4,4) W2C^T flattened row-major
    proj: np.ndarray   # 16 floats, full projection, same convention
    campos: np.ndarray  # 3 floats

    @property
    def grid(self):
        return (self.W + 15) // 16, (self.H + 15) // 16

    @property
    def tiles(self):
        gx, gy = self.grid
        return gx * gy


@dataclass
class FwdState:
    """Everything the forward produced; what a backward and the parity tests need."""

    P: int
    D: int
    M: int
    radii: np.ndarray
    xy: np.ndarray
    depths: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    offsets: np.ndarray
    num_rendered: int
    keys_unsorted: np.ndarray
    vals_unsorted: np.ndarray
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    color: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray
    alpha: np.ndarray
    depth: np.ndarray
    inputs: dict = field(default_factory=dict)


def forward(cam: Camera, means3D, opacities, bg, shs=None, colors_precomp=None, scales=None,
            rotations=None, cov3D_precomp=None, sh_degree: int = 0,
            scale_modifier: float = 1.0) -> FwdState:
    """Full forward: preprocess -> scan -> duplicateWithKeys -> sort -> ranges -> render
    (SURVEY.md A.1 orchestration)."""
    L_ = lib()
    means3D = _f32(means3D)
    P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    shs = _f32(shs)
    colors_precomp = _f32(colors_precomp)
    scales = _f32(scales)
    rotations = _f32(rotations)
    cov3D_precomp = _f32(cov3D_precomp)
    bg = _f32(bg)
    view, proj, campos = _f32(cam.view).reshape(-1), _f32(cam.proj).reshape(-1), _f32(cam.campos)
    M = shs.shape[1] if shs is not None else 0
    W, H = cam.W, cam.H

    radii = np.zeros(P, np.int32)
    xy = np.zeros((P, 2), np.float32)
    depths = np.zeros(P, np.float32)
    cov3D = np.zeros((P, 6), np.float32)
    conic_opacity = np.zeros((P, 4), np.float32)
    rgb = np.zeros((P, 3), np.float32)
    clamped = np.zeros((P, 3), np.uint8)
    tiles_touched = np.zeros(P, np.uint32)
    L_.ro_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D, _f32p), _p(scales, _f32p),
        C.c_float(scale_modifier), _p(rotations, _f32p), _p(opacities, _f32p), _p(shs, _f32p),
        _p(colors_precomp, _f32p), _p(cov3D_precomp, _f32p), _p(view, _f32p), _p(proj, _f32p),
        _p(campos, _f32p), C.c_int(W), C.c_int(H), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy),
        _p(radii, _i32p), _p(xy, _f32p), _p(depths, _f32p), _p(cov3D, _f32p),
        _p(conic_opacity, _f32p), _p(rgb, _f32p), _p(clamped, _u8p), _p(tiles_touched, _u32p))
    offsets = np.zeros(P, np.uint32)
    L = int(L_.ro_scan(C.c_int(P), _p(tiles_touched, _u32p), _p(offsets, _u32p))) if P else 0
    keys_u = np.zeros(max(L, 1), np.uint64)
    vals_u = np.zeros(max(L, 1), np.uint32)
    L_.ro_duplicate_with_keys(C.c_int(P), C.c_int(W), C.c_int(H), _p(xy, _f32p),
                              _p(depths, _f32p), _p(offsets, _u32p), _p(radii, _i32p),
                              _p(keys_u, _u64p), _p(vals_u, _u32p))
    keys = keys_u.copy()
    vals = vals_u.copy()
    tk = np.zeros_like(keys)
    tv = np.zeros_like(vals)
    end_bit = 32 + higher_msb(cam.tiles)
    L_.ro_sort_pairs(C.c_int64(L), _p(keys, _u64p), _p(vals, _u32p), _p(tk, _u64p),
                     _p(tv, _u32p), C.c_int(end_bit))
    ranges = np.zeros((cam.tiles, 2), np.uint32)
    L_.ro_tile_ranges(C.c_int64(L), _p(keys, _u64p), C.c_int(cam.tiles), _p(ranges, _u32p))
    color = np.zeros((3, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    alpha = np.zeros((H, W), np.float32)
    depth = np.zeros((H, W), np.float32)
    L_.ro_render_fwd(C.c_int(W), C.c_int(H), _p(ranges, _u32p), _p(vals, _u32p), _p(xy, _f32p),
                     _p(conic_opacity, _f32p), _p(rgb, _f32p), _p(depths, _f32p), _p(bg, _f32p),
                     _p(color, _f32p), _p(final_T, _f32p), _p(n_contrib, _u32p),
                     _p(alpha, _f32p), _p(depth, _f32p))
    return FwdState(
        P=P, D=sh_degree, M=M, radii=radii, xy=xy, depths=depths, cov3D=cov3D,
        conic_opacity=conic_opacity, rgb=rgb, clamped=clamped, tiles_touched=tiles_touched,
        offsets=offsets, num_rendered=L, keys_unsorted=keys_u[:L], vals_unsorted=vals_u[:L],
        keys=keys[:L], point_list=vals[:L] if L else vals[:0], ranges=ranges, color=color,
        final_T=final_T, n_contrib=n_contrib, alpha=alpha, depth=depth,
        inputs=dict(cam=cam, means3D=means3D, opacities=opacities, bg=bg, shs=shs,
                    colors_precomp=colors_precomp, scales=scales, rotations=rotations,
                    cov3D_precomp=cov3D_precomp, scale_modifier=scale_modifier))


def backward(st: FwdState, dL_dcolor_img) -> dict:
    """Full backward for dL/d(out_color) (3,H,W): render bwd -> cov2D/preprocess bwd.
    Returns the gradients in the upstream autograd order plus the raw per-Gaussian sums."""
    L_ = lib()
    inp = st.inputs
    cam: Camera = inp["cam"]
    W, H, P = cam.W, cam.H, st.P
    dpix = _f32(dL_dcolor_img)
    dmean2D = np.zeros((P, 2), np.float32)
    dconic = np.zeros((P, 3), np.float32)
    dopac = np.zeros(P, np.float32)
    dcolor = np.zeros((P, 3), np.float32)
    plist = np.ascontiguousarray(st.point_list) if st.num_rendered else np.zeros(1, np.uint32)
    L_.ro_render_bwd(C.c_int(W), C.c_int(H), _p(st.ranges, _u32p), _p(plist, _u32p),
                     _p(st.xy, _f32p), _p(st.conic_opacity, _f32p), _p(st.rgb, _f32p),
                     _p(inp["bg"], _f32p), _p(st.final_T, _f32p), _p(st.n_contrib, _u32p),
                     _p(dpix, _f32p), _p(dmean2D, _f32p), _p(dconic, _f32p), _p(dopac, _f32p),
                     _p(dcolor, _f32p))
    dmeans3D = np.zeros((P, 3), np.float32)
    dscales = np.zeros((P, 3), np.float32)
    drots = np.zeros((P, 4), np.float32)
    dsh = np.zeros((P, st.M, 3), np.float32) if inp["shs"] is not None else None
    dcov3D = np.zeros((P, 6), np.float32)
    view, proj = _f32(cam.view).reshape(-1), _f32(cam.proj).reshape(-1)
    L_.ro_preprocess_bwd(
        C.c_int(P), C.c_int(st.D), C.c_int(st.M), _p(inp["means3D"], _f32p),
        _p(inp["scales"], _f32p), C.c_float(inp["scale_modifier"]), _p(inp["rotations"], _f32p),
        _p(inp["shs"], _f32p), _p(inp["cov3D_precomp"], _f32p), _p(view, _f32p), _p(proj, _f32p),
        _p(_f32(cam.campos), _f32p), C.c_int(W), C.c_int(H), C.c_float(cam.tanfovx),
        C.c_float(cam.tanfovy), _p(st.radii, _i32p), _p(st.clamped, _u8p), _p(dmean2D, _f32p),
        _p(dconic, _f32p), _p(dcolor, _f32p), _p(dmeans3D, _f32p), _p(dscales, _f32p),
        _p(drots, _f32p), _p(dsh, _f32p), _p(dcov3D, _f32p))
    dmeans2D = np.zeros((P, 3), np.float32)
    dmeans2D[:, :2] = dmean2D
    return dict(means3D=dmeans3D, means2D=dmeans2D, sh=dsh, colors_precomp=dcolor,
                opacities=dopac.reshape(P, 1), scales=dscales, rotations=drots,
                cov3Ds_precomp=dcov3D, conic=dconic)


def render_mask(st: FwdState, tile: int) -> np.ndarray:
    """(256, K) uint8 decision mask of one tile (see ro_render_mask)."""
    cam: Camera = st.inputs["cam"]
    r0, r1 = int(st.ranges[tile, 0]), int(st.ranges[tile, 1])
    K = r1 - r0
    mask = np.zeros((256, max(K, 1)), np.uint8)
    if K > 0:
        lib().ro_render_mask(C.c_int(cam.W), C.c_int(cam.H), _p(st.ranges, _u32p),
                             _p(np.ascontiguousarray(st.point_list), _u32p), _p(st.xy, _f32p),
                             _p(st.conic_opacity, _f32p), _p(st.n_contrib, _u32p), C.c_int(tile),
                             _p(mask, _u8p))
    return mask[:, :K]


def mark_visible(means3D, view) -> np.ndarray:
    means3D = _f32(means3D)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().ro_mark_visible(C.c_int(means3D.shape[0]), _p(means3D, _f32p),
                          _p(_f32(view).reshape(-1), _f32p), _p(out, _u8p))
    return out.astype(bool)

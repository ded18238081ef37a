"""Host side of the rasterizer: torch.autograd wrapper over the C ABI (include/sings_b200.h).

Mirrors the Python surface of the reference's third-party rasterizer -- the classic
`diff_gaussian_rasterization` API that SinGS is written against
(/root/reference/sings/rec/renderer/gs_renderer_single.py:6-9, 69-95;
gs_renderer_multiple.py:6-9, 95-121): `GaussianRasterizationSettings` with exactly those 12
fields, `GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs | colors_precomp,
scales + rotations | cov3D_precomp) -> (color (3,H,W), radii (P,))`, `markVisible`, and the
gradient order `(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
cov3Ds_precomp, None)`.

Differences by design (B200-first, results identical):
  * the forward never blocks on a device->host copy to size buffers: the (tile, Gaussian)
    pair list is written into a capacity-sized buffer; the pair count and an overflow flag
    come back through pinned memory.  Default ("checked") mode waits for that flag once per
    forward and transparently re-runs with a larger buffer if it overflowed; `set_async(True)`
    defers the check to `check_pending()` / the next forward (no host sync at all).
  * scratch is torch-owned (`torch.empty(uint8)`), sized by `sgs_raster_sizes`; the library
    never allocates.
  * alpha (= 1 - final transmittance) and expected depth are available from
    `GaussianRasterizer.forward_aux(...)` without changing the 2-tuple contract.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import threading
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib

_ASYNC = os.environ.get("SGS_ASYNC", "0") == "1"
_cap_hint: dict = {}           # device index -> pair-list capacity learned from earlier frames
_pinned: dict = {}             # device index -> (pinned int32 [slots,2], next slot)
_pending: list = []            # async mode: (event, pinned row, L_cap) not yet checked
_SLOTS = 256
last_num_rendered = 0          # informational: pair count of the last checked forward
_state_lock = threading.Lock() # host threads rendering concurrently (each on its own stream) share the tables above


def set_async(flag: bool) -> None:
    """Async mode: forward() performs no host synchronisation; a pair-list overflow is raised
    by check_pending() (called at the start of every forward) instead of being repaired."""
    global _ASYNC
    _ASYNC = bool(flag)


def _pinned_row(dev: int):
    with _state_lock:
        buf, nxt = _pinned.get(dev, (None, 0))
        if buf is None:
            buf = torch.zeros(_SLOTS, 2, dtype=torch.int32).pin_memory()
        _pinned[dev] = (buf, (nxt + 1) % _SLOTS)
    return buf[nxt]


def _raise_cap_hint(dev: int, L: int) -> int:
    with _state_lock:
        _cap_hint[dev] = max(_cap_hint.get(dev, 0), int(L * 1.3) + 4096)
        return _cap_hint[dev]


def check_pending(block: bool = False) -> None:
    """Async mode: examine finished forwards; raise if one overflowed its pair-list capacity."""
    global last_num_rendered
    with _state_lock:
        todo = list(_pending)
        _pending.clear()
    keep = []
    try:
        while todo:
            ev, row, cap, dev = todo.pop(0)
            if block:
                ev.synchronize()
            if ev.query():
                L, ovf = int(row[0]), int(row[1])
                last_num_rendered = L
                _raise_cap_hint(dev, L)
                if ovf:
                    todo.clear()
                    keep.clear()
                    raise _lib.SgsError(
                        f"rasterizer pair list overflowed (needed {L}, capacity {cap}) in async mode; "
                        "the frame is incomplete. Capacity has been raised; re-render the frame.")
            else:
                keep.append((ev, row, cap, dev))
    finally:
        with _state_lock:
            _pending[:0] = keep + todo


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool = False     # accepted for forward compatibility; must stay False


def _f32c(t: Optional[torch.Tensor], name: str, dev) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if t.device != dev:
        raise ValueError(f"{name} must live on {dev}, got {t.device}")
    if t.dtype is torch.float32 and t.is_contiguous():      # the usual case: no new tensor
        return t
    return t.float().contiguous()


@functools.lru_cache(maxsize=64)
def _sizes(P, W, H, L_cap):
    """(geom, binning, img, acc) scratch bytes; one C call per distinct shape (cached)."""
    s = [C.c_size_t() for _ in range(4)]
    _lib.check(_lib.lib().sgs_raster_sizes(P, W, H, L_cap, *[C.byref(x) for x in s]), "sgs_raster_sizes")
    return tuple(int(x.value) for x in s)


def _align256(n: int) -> int:
    return (n + 255) // 256 * 256


def layout_info(P, W, H, L_cap) -> dict:
    """Offsets of the inspectable arrays inside the scratch buffers (parity tests)."""
    info = (C.c_longlong * 16)()
    _lib.check(_lib.lib().sgs_raster_layout_info(P, W, H, L_cap, info), "sgs_raster_layout_info")
    keys = ["counters", "keys_unsorted", "vals_unsorted", "keys_sorted", "vals_sorted", "ranges",
            "final_T", "n_contrib", "tiles", "end_bit", "passes", "rec_floats"]
    return {k: int(info[i]) for i, k in enumerate(keys)}


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, want_aux):
        L_ = _lib.lib()
        rs = raster_settings
        if getattr(rs, "antialiasing", False):
            raise NotImplementedError("antialiasing=True is not part of the classic API SinGS uses")
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise ValueError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise _lib.SgsError("sings_b200 rasterizer needs CUDA tensors (no CPU fallback)")
        dev = means3D.device
        if dev.index is not None and dev.index != torch.cuda.current_device():
            # the library launches on `dev`'s current stream: make `dev` the current device for the call
            with torch.cuda.device(dev):
                return _RasterizeGaussians.forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales,
                                                   rotations, cov3Ds_precomp, raster_settings, want_aux)
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        m3 = _f32c(means3D, "means3D", dev)
        if m3 is None:
            m3 = torch.zeros(0, 3, device=dev)
        shc = _f32c(sh, "shs", dev)
        col = _f32c(colors_precomp, "colors_precomp", dev)
        opa = _f32c(opacities, "opacities", dev)
        sca = _f32c(scales, "scales", dev)
        rot = _f32c(rotations, "rotations", dev)
        cov = _f32c(cov3Ds_precomp, "cov3D_precomp", dev)
        bg = _f32c(rs.bg, "bg", dev)
        view = _f32c(rs.viewmatrix, "viewmatrix", dev)
        proj = _f32c(rs.projmatrix, "projmatrix", dev)
        campos = _f32c(rs.campos, "campos", dev)
        M = shc.shape[1] if shc is not None else 0
        D = int(rs.sh_degree)
        if P > 0:
            if (shc is None) == (col is None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if (cov is None) == (sca is None or rot is None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
            if opa is None or opa.numel() != P:
                raise ValueError("opacities must have one value per Gaussian")
        st = _lib.raw_stream(dev)
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        if _ASYNC:
            check_pending()
        L_cap = max(_cap_hint.get(di, 0), 4 * P, 1 << 16)
        color = torch.empty(3, H, W, device=dev, dtype=torch.float32)
        radii = torch.empty(P, device=dev, dtype=torch.int32)
        alpha = torch.empty(H, W, device=dev, dtype=torch.float32) if want_aux else None
        depth = torch.empty(H, W, device=dev, dtype=torch.float32) if want_aux else None
        global last_num_rendered
        needs_grad = any(x is not None and x.requires_grad for x in
                         (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
        while True:
            # ONE allocation for the frame's state (geometry records | binning | image state | backward
            # accumulator) and ONE kernel that zeroes what forward and backward need zeroed: no memset
            # node sits between two kernels (each costs them their overlapped launch), and the clear
            # kernel in front lets the SH rows be fetched ahead of the dependency wait (EARLY_PARAMS:
            # whatever produced `shs` finished before our clear kernel released its dependents).
            gb, bb, ib, ab = _sizes(P, W, H, L_cap)
            o1, o2, o3 = _align256(gb), _align256(gb) + _align256(bb), _align256(gb) + _align256(bb) + _align256(ib)
            scratch = torch.empty(o3 + (ab if needs_grad else 0), device=dev, dtype=torch.uint8)
            geom, binning, img = scratch[:gb], scratch[o1:o1 + bb], scratch[o2:o2 + ib]
            acc = scratch[o3:o3 + ab] if needs_grad else None
            row = _pinned_row(di)
            flags = int(bool(rs.debug))
            if not needs_grad:
                flags |= _lib.FLAG_FORWARD_ONLY        # no backward can follow: skip what the forward leaves for it
            if P > 0:
                _lib.check(L_.sgs_raster_clear(P, W, H, L_cap, _lib.ptr(binning), _lib.ptr(acc), None, 0,
                                               st), "sgs_raster_clear")
                flags |= _lib.FLAG_PRECLEARED | _lib.FLAG_EARLY_PARAMS
            try:
                rc = L_.sgs_raster_forward(
                    P, D, M, W, H, _lib.ptr(bg), _lib.ptr(m3), _lib.ptr(col), _lib.ptr(opa),
                    _lib.ptr(sca), float(rs.scale_modifier), _lib.ptr(rot), _lib.ptr(cov),
                    _lib.ptr(view), _lib.ptr(proj), _lib.ptr(campos), float(rs.tanfovx),
                    float(rs.tanfovy), _lib.ptr(shc), int(bool(rs.prefiltered)), L_cap,
                    _lib.ptr(geom), _lib.ptr(binning), _lib.ptr(img), _lib.ptr(color),
                    _lib.ptr(radii), _lib.ptr(alpha), _lib.ptr(depth), row.data_ptr(),
                    st, flags, None)
                _lib.check(rc, "sgs_raster_forward")
            except Exception:
                if rs.debug:
                    torch.save((means3D, sh, colors_precomp, opacities, scales, rotations,
                                cov3Ds_precomp, tuple(rs)), "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
            ev = torch.cuda.Event()
            ev.record()                        # current stream of the current device == `dev` (see the guard above)
            ctx.fwd_check = None
            if _ASYNC:
                with _state_lock:
                    _pending.append((ev, row, L_cap, di))
                ctx.fwd_check = (ev, row, L_cap)
                break
            ev.synchronize()
            L, ovf = int(row[0]), int(row[1])
            last_num_rendered = L
            hint = _raise_cap_hint(di, L)
            if not ovf:
                break
            L_cap = hint
        ctx.raster_settings = rs
        ctx.L_cap = L_cap
        ctx.dims = (P, D, M, W, H)
        ctx.has = (shc is not None, col is not None, cov is not None)
        ctx.opac_shape = tuple(opacities.shape) if opacities is not None else (P, 1)
        ctx.acc_clean = acc is not None and P > 0
        ctx.save_for_backward(m3, shc, col, sca, rot, cov, radii, geom, binning, img, bg, view,
                              proj, campos, acc)
        ctx.mark_non_differentiable(radii)
        if want_aux:
            ctx.mark_non_differentiable(alpha, depth)
            return color, radii, alpha, depth
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, *_unused):
        L_ = _lib.lib()
        rs = ctx.raster_settings
        P, D, M, W, H = ctx.dims
        (m3, shc, col, sca, rot, cov, radii, geom, binning, img, bg, view, proj,
         campos, acc) = ctx.saved_tensors
        dev = m3.device
        if dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return _RasterizeGaussians.backward(ctx, grad_out_color, *_unused)
        if ctx.fwd_check is not None:
            # async mode: this frame's forward has not been examined yet -- gradients of a truncated pair
            # list must not reach the optimizer (the event is long complete by now: no real wait)
            ev, row, cap = ctx.fwd_check
            ev.synchronize()
            if int(row[1]):
                raise _lib.SgsError(f"rasterizer pair list overflowed in the forward of this frame (needed "
                                    f"{int(row[0])}, capacity {cap}); its gradients are not valid. Re-render the frame.")
        g = grad_out_color
        if g.dtype != torch.float32:
            g = g.float()
        g = g.contiguous()
        st = _lib.raw_stream(dev)
        # the per-Gaussian gradients as views of one allocation (rots and sh first: 16-byte aligned)
        has_sh, has_col, has_cov = ctx.has
        widths = [4, 3 * M if has_sh else 0, 3, 3, 3, 1, 6, 3]
        flat = torch.empty(P * sum(widths), device=dev, dtype=torch.float32)
        views, o = [], 0
        for w_ in widths:
            views.append(flat[o:o + P * w_].view(P, w_) if w_ else None)
            o += P * w_
        d_rots, d_sh, d_means3D, d_means2D, d_colors, d_opac, d_cov, d_scales = views
        if d_sh is not None:
            d_sh = d_sh.view(P, M, 3)
        flags = int(bool(rs.debug))
        if ctx.acc_clean:
            flags |= _lib.FLAG_PRECLEARED      # cleared together with the forward's state, by one kernel
            ctx.acc_clean = False              # a second backward of the same graph clears it again (memset)
        try:
            rc = L_.sgs_raster_backward(
                P, D, M, W, H, _lib.ptr(bg), _lib.ptr(m3), _lib.ptr(col), _lib.ptr(sca),
                float(rs.scale_modifier), _lib.ptr(rot), _lib.ptr(cov), _lib.ptr(view),
                _lib.ptr(proj), _lib.ptr(campos), float(rs.tanfovx), float(rs.tanfovy),
                _lib.ptr(shc), _lib.ptr(radii), _lib.ptr(g), ctx.L_cap, _lib.ptr(geom),
                _lib.ptr(binning), _lib.ptr(img), _lib.ptr(acc), _lib.ptr(d_means3D),
                _lib.ptr(d_means2D), _lib.ptr(d_colors), _lib.ptr(d_opac), _lib.ptr(d_cov),
                _lib.ptr(d_sh), _lib.ptr(d_scales), _lib.ptr(d_rots), None, None, None,
                st, flags, None)
            _lib.check(rc, "sgs_raster_backward")
        except Exception:
            if rs.debug:
                torch.save((m3, radii, col, sca, rot, cov, g, shc, tuple(rs)), "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise
        return (d_means3D, d_means2D, d_sh if has_sh else None, d_colors if has_col else None,
                d_opac.reshape(ctx.opac_shape), None if has_cov else d_scales, None if has_cov else d_rots,
                d_cov if has_cov else None, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, want_aux=False):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings, want_aux)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of Gaussians in front of the near plane (view z > 0.2)."""
        with torch.no_grad():
            rs = self.raster_settings
            pos = positions.detach().float().contiguous()
            view = rs.viewmatrix.float().contiguous()
            out = torch.empty(pos.shape[0], device=pos.device, dtype=torch.uint8)
            _lib.check(_lib.lib().sgs_mark_visible(
                pos.shape[0], pos.data_ptr(), view.data_ptr(), out.data_ptr(),
                torch.cuda.current_stream(pos.device).cuda_stream), "sgs_mark_visible")
            return out.bool()

    def _check(self, shs, colors_precomp, scales, rotations, cov3D_precomp):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        self._check(shs, colors_precomp, scales, rotations, cov3D_precomp)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, self.raster_settings, False)

    def forward_aux(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                    rotations=None, cov3D_precomp=None):
        """Like forward(), plus alpha (H,W) = 1 - final_T and depth (H,W) = sum alpha_i T_i z_i
        (both non-differentiable)."""
        self._check(shs, colors_precomp, scales, rotations, cov3D_precomp)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, self.raster_settings, True)

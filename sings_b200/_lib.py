"""ctypes binding of libsings_b200.so (C ABI: include/sings_b200.h).

There is no CPU path and no fallback: if the CUDA library is missing or a call fails, the
caller gets an exception.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# SGS_LIB_PATH: an alternative build of the same library (A/B variants built by tools/variants.sh)
LIB_PATH = os.environ.get("SGS_LIB_PATH") or os.path.join(_HERE, "lib", "libsings_b200.so")
CSRC = os.path.join(_HERE, "csrc")
_lib = None

c_f32p = C.c_void_p   # device pointers travel as integers
_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

class DeformArgs(C.Structure):
    """sgs_deform_args (include/sings_b200.h): the deformer side of the fused frame calls."""
    _fields_ = [("N", _i), ("J", _i), ("K", _i), ("rot6d", _i)] + [
        (name, _vp) for name in (
            "pose", "rest", "parents", "inv_A_t2cano", "xyz_canon", "scales", "rot_canon", "wq", "iq",
            "smpl_scale", "transl", "A", "G", "xyz", "rotq", "scales_out", "d_xyz_canon", "d_rot_canon",
            "d_scales", "d_A", "d_transl", "d_pose")]


_SIGNATURES = {
    "sgs_version": (C.c_int, []),
    "sgs_error_string": (C.c_char_p, [_i]),
    "sgs_timing_create": (_i, [_i, C.POINTER(_vp)]),
    "sgs_timing_destroy": (_i, [_vp]),
    "sgs_timing_record": (_i, [_vp, _i, _vp]),
    "sgs_timing_set_mask": (_i, [_vp, C.c_uint]),
    "sgs_timing_elapsed_ms": (_i, [_vp, _i, _i, C.POINTER(_f)]),
    "sgs_graph_begin": (_i, [_vp]),
    "sgs_graph_end": (_i, [_vp, C.POINTER(_vp)]),
    "sgs_graph_launch": (_i, [_vp, _vp]),
    "sgs_graph_destroy": (_i, [_vp]),
    "sgs_raster_sizes": (_i, [_i, _i, _i, _ll, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "sgs_raster_layout_info": (_i, [_i, _i, _i, _ll, C.POINTER(_ll)]),
    "sgs_raster_clear": (_i, [_i, _i, _i, _ll, _vp, _vp, _vp, _sz, _vp]),
    "sgs_raster_forward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp,
                                _vp, _f, _f, _vp, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _i, _vp]),
    "sgs_raster_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                 _f, _f, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sgs_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp]),
    "sgs_densify_stats": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_fold_stats": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_frame_to_u8": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "sgs_hexplane_fwd": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "sgs_hexplane_bwd": (_i, [_i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sgs_knn_scratch_bytes": (_sz, [_i]),
    "sgs_knn_mean_dist": (_i, [_i, _vp, _i, _vp, _sz, _vp, _vp, _vp, _vp]),
    "sgs_image_loss_scratch_floats": (_sz, [_i, _i]),
    "sgs_image_loss_fwd": (_i, [_i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp]),
    "sgs_image_loss_bwd": (_i, [_i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "sgs_laplacian_loss_fwd": (_i, [_i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "sgs_laplacian_loss_bwd": (_i, [_i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "sgs_l2norm_fwd": (_i, [_i, _vp, _vp, _i, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp]),
    "sgs_l2norm_bwd": (_i, [_i, _vp, _vp, _i, _i, _vp, _f, _f, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "sgs_sort_scratch_bytes": (_sz, [_ll]),
    "sgs_sort_pairs_u64": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _ll, _i, C.POINTER(_i), _vp]),
    "sgs_pose_to_A": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sgs_pose_to_A_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "sgs_lbs_fwd": (_i, [_i, _i, _i] + [_vp] * 15),
    "sgs_lbs_bwd": (_i, [_i, _i, _i] + [_vp] * 21),
    "sgs_pose_lbs_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i] + [_vp] * 12),
    "sgs_lbs_fwd_rot6d": (_i, [_i, _i, _i] + [_vp] * 15),
    "sgs_lbs_bwd_rot6d": (_i, [_i, _i, _i] + [_vp] * 21),
    "sgs_lbs_packed_bytes": (_sz, [_i, _i]),
    "sgs_lbs_pack_weights": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "sgs_avatar_forward": (_i, [C.POINTER(DeformArgs), _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _f, _f, _vp,
                                _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sgs_avatar_backward": (_i, [C.POINTER(DeformArgs), _i, _i, _i, _i, _vp, _f, _vp, _vp, _vp, _f, _f, _vp, _vp,
                                 _vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sgs_rot6d_to_matrix": (_i, [_vp, _i, _vp, _vp]),
    "sgs_rot6d_to_matrix_bwd": (_i, [_vp, _vp, _i, _vp, _vp]),
    "sgs_rot6d_to_axis_angle": (_i, [_vp, _i, _vp, _vp]),
    "sgs_rot6d_to_axis_angle_bwd": (_i, [_vp, _vp, _i, _vp, _vp]),
}

EXPORTS = tuple(_SIGNATURES)
FLAG_SYNC_CHECK, FLAG_PRECLEARED, FLAG_EARLY_PARAMS, FLAG_FORWARD_ONLY = 1, 2, 4, 8   # bits of the rasterizer entry points' `debug` argument


class SgsError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (sings_b200/csrc/Makefile)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    out = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise SgsError("building libsings_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SgsError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sings_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().sgs_error_string(code).decode()
        raise SgsError(f"{what}: {msg} (code {code})" if what else f"{msg} (code {code})")


def raw_stream(dev) -> int:
    """cudaStream_t of torch's current stream on `dev` (the raw getter: ~1 us instead of the ~12 us of
    torch.cuda.current_stream(), which the drop-in path would otherwise pay six times per step)."""
    import torch
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


def ptr(t) -> int | None:
    """data_ptr of a CUDA tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()

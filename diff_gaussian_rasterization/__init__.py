"""Drop-in for the `diff_gaussian_rasterization` package SinGS imports
(/root/reference/sings/rec/renderer/gs_renderer_single.py:6-9,
 /root/reference/sings/rec/renderer/gs_renderer_multiple.py:6-9; installed upstream by
 /root/reference/install_all.sh:22).

Same names, same 12-field settings tuple, same call signature, same gradient order -- backed
by the sm_100a kernels of sings_b200 (no upstream code, no CPU fallback).
"""
from sings_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    set_async,
    check_pending,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]

"""sings_b200 -- B200 (sm_100a) implementation of SinGS's per-frame avatar hot path:
SMPL linear-blend-skinning of every Gaussian followed by the 3DGS differentiable tile
rasterizer, forward and backward, behind the reference's own call signatures.

    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from sings_b200.deform import deform_gaussians, lbs_extra, pose_to_A

Everything runs in hand-written CUDA kernels reached through the C ABI of
sings_b200/lib/libsings_b200.so (include/sings_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"

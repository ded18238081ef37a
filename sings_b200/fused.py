"""One differentiable call for the whole per-frame path: pose -> A -> deform -> rasterize.

The drop-in modules (`sings_b200.deform` + `diff_gaussian_rasterization.GaussianRasterizer`) keep
the reference's call structure -- three autograd Functions, per-call allocation -- and are bound
by the host (torch.autograd + Python, ~0.8 ms per step at any GPU speed; profiles/
r02_dropin_host_profile.txt).  `AvatarRenderer` is the opt-in alternative for a trainer that is
willing to change five lines: the same computation as ONE autograd Function
over the fused kernels and preallocated buffers of `sings_b200.step.AvatarStep`
(sgs_avatar_forward / sgs_avatar_backward), still an ordinary differentiable torch call:

    renderer = AvatarRenderer(xyz_canon, rotmat_canon, scales, opacity, shs, lbs_weights,
                              rest_joints, parents, inv_A_t2cano, H, W, sh_degree)
    image, radii = renderer(pose, transl, viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy)
    loss(image).backward()          # .grad on the parameter tensors, on pose and on transl

Replaces, in one call, sings_hybrid.py:370-428 (pose conversion excluded) + gs_renderer_single.py:
48-101.  The parameter tensors are read in place (float32, contiguous: no copies), so optimizer
steps are seen by the next call; after a densification (N changes) build a new AvatarRenderer.
There is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import SgsError
from .step import AvatarStep, FrameInputs


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, pose, transl, smpl_scale, xyz, rot, scales, opacity, shs, cam):
        st: AvatarStep = owner._step
        viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy = cam
        fr = FrameInputs(pose=pose.detach().reshape(-1, 3), transl=transl.detach().reshape(3), viewmatrix=viewmatrix,
                         projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=float(tanfovx), tanfovy=float(tanfovy),
                         smpl_scale=None if smpl_scale is None else smpl_scale.detach().reshape(1))
        needs_grad = any(t is not None and t.requires_grad for t in (pose, transl, xyz, rot, scales, opacity, shs))
        st.forward_only = not needs_grad
        owner._slot = (owner._slot + 1) % st.COUNTER_SLOTS
        while True:
            img = st.forward(fr, slot=owner._slot)
            if not owner.sync_check:
                break
            torch.cuda.current_stream(st.dev).synchronize()      # like the reference's num_rendered read-back
            try:
                st.check_capacity()
                break
            except SgsError:
                continue                                          # the pair list grew: render again
        ctx.owner = owner
        ctx.gen = owner._calls = owner._calls + 1
        ctx.mark_non_differentiable(st.radii)
        return img.clone(), st.radii.clone()

    @staticmethod
    def backward(ctx, dL_dimage, _dradii):
        owner = ctx.owner
        st: AvatarStep = owner._step
        if ctx.gen != owner._calls:
            raise SgsError("AvatarRenderer: backward of a frame that is not the renderer's latest forward "
                           "(the buffers hold one frame: call backward before the next forward)")
        st.backward(dL_dimage.contiguous().float(), stats=owner.densify_stats)
        c = (lambda t: t.clone()) if owner.clone_grads else (lambda t: t)
        g = lambda need, t: c(t) if (need and t is not None) else None
        n = ctx.needs_input_grad
        return (None, g(n[1], st.d_pose.view(-1, 3)), g(n[2], st.d_transl.view(3)), None,
                g(n[4], st.d_xyz_canon), g(n[5], st.d_rot_canon), g(n[6], st.d_scales), g(n[7], st.d_opacity),
                g(n[8], st.d_shs), None)


class AvatarRenderer:
    """See the module docstring.  sync_check=True (default) examines the pair-list overflow flag
    after every forward (one stream synchronisation per frame, what the reference's rasterizer does
    too) and re-renders transparently; sync_check=False never blocks the host -- call
    `check()` every few frames (raises SgsError if a frame overflowed and had to be dropped).
    clone_grads=False hands autograd views of the persistent gradient bucket (valid until the next
    backward; fine for optimizers that consume .grad right away).  densify_stats=True accumulates
    xyz_gradient_accum / denom / max_radii2D of every backward in `step` (sings_hybrid.py:1013-1015)."""

    def __init__(self, xyz_canon, rotmat_canon, scales, opacity, shs, lbs_weights, rest_joints, parents,
                 inv_A_t2cano, H: int, W: int, sh_degree: int, sync_check: bool = True, clone_grads: bool = True,
                 densify_stats: bool = True):
        for name, t in (("xyz_canon", xyz_canon), ("rotmat_canon", rotmat_canon), ("scales", scales),
                        ("opacity", opacity), ("shs", shs)):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise SgsError(f"AvatarRenderer reads {name} in place: it must be float32 and contiguous")
        self.params = (xyz_canon, rotmat_canon, scales, opacity, shs)
        self._step = AvatarStep(xyz_canon, rotmat_canon, scales, opacity, shs, lbs_weights, rest_joints, parents,
                                inv_A_t2cano, H, W, sh_degree)
        self.sync_check, self.clone_grads, self.densify_stats = bool(sync_check), bool(clone_grads), bool(densify_stats)
        self._slot, self._calls = 0, 0

    @property
    def step(self) -> AvatarStep:
        return self._step

    def __call__(self, pose, transl, viewmatrix, projmatrix, campos, bg, tanfovx: float, tanfovy: float,
                 smpl_scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """pose (J, 3) axis-angle, transl (3,), camera as in GaussianRasterizationSettings
        (viewmatrix / projmatrix transposed like SinGS passes them).  Returns (image (3, H, W), radii (N,))."""
        xyz, rot, scales, opacity, shs = self.params
        cam = (viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy)
        return _Render.apply(self, pose, transl, smpl_scale, xyz, rot, scales, opacity, shs, cam)

    def check(self) -> int:
        """sync_check=False: wait for the device and examine the overflow flags of the frames since the
        last check.  Returns the largest pair count; raises SgsError (after growing the list) if a
        frame overflowed -- its image and gradients were garbage."""
        torch.cuda.current_stream(self._step.dev).synchronize()
        return self._step.check_capacity()

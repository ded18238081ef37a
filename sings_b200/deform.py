"""Host side of the deformer: torch.autograd wrappers over the fused LBS kernels.

Mirrors the reference's deform segment (/root/reference/sings/rec/models/sings_hybrid.py:
398-428 `forward`, :525-552 `forward_chunk`) and its helper signature
`lbs_extra(A, v_shaped, posedirs, lbs_weights, pose, disable_posedirs, pose2rot)
 -> (verts, A, T, v_posed, v_shaped)` (/root/reference/sings/rec/utils/body_model/lbs.py:16-74).

`deform_gaussians` is the fused fast path: one kernel produces xyz / rotq / scales for all B
frames; T is materialised only when asked for.  Gradients flow to xyz_canon, rotmat_canon,
scales, A (and through `pose_to_A` to the pose), smpl_scale and transl.  `lbs_weights` and the
external similarity `ext_tfs` receive no gradient (the reference never optimises them:
sings_hybrid.py:724 registers the weights as a buffer, ext_tfs is only used under no_grad in
gs_trainer.py:629).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.SgsError(f"{what}: sings_b200 needs CUDA tensors (no CPU fallback)")


class _PoseToA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, rest, parents, inv_A):
        _need_cuda(pose, "pose_to_A")
        pose_c, rest_c = _c(pose), _c(rest)
        inv_c = _c(inv_A)
        par = parents.to(device=pose.device, dtype=torch.int32).contiguous()
        B, J = pose_c.shape[0], pose_c.shape[1]
        A = torch.empty(B, J, 4, 4, device=pose.device, dtype=torch.float32)
        G = torch.empty(B, J, 12, device=pose.device, dtype=torch.float32)
        st = _lib.raw_stream(pose.device)
        _lib.check(_lib.lib().sgs_pose_to_A(pose_c.data_ptr(), rest_c.data_ptr(), par.data_ptr(),
                                            _lib.ptr(inv_c), B, J, A.data_ptr(), G.data_ptr(), st),
                   "sgs_pose_to_A")
        ctx.save_for_backward(pose_c, rest_c, par, inv_c, G)
        return A

    @staticmethod
    def backward(ctx, dA):
        pose_c, rest_c, par, inv_c, G = ctx.saved_tensors
        B, J = pose_c.shape[0], pose_c.shape[1]
        d_pose = torch.empty_like(pose_c)
        st = _lib.raw_stream(pose_c.device)
        _lib.check(_lib.lib().sgs_pose_to_A_bwd(pose_c.data_ptr(), rest_c.data_ptr(), par.data_ptr(),
                                                _lib.ptr(inv_c), G.data_ptr(), _c(dA).data_ptr(), B, J,
                                                d_pose.data_ptr(), st), "sgs_pose_to_A_bwd")
        return d_pose, None, None, None


def pose_to_A(pose: torch.Tensor, rest_joints: torch.Tensor, parents: torch.Tensor,
              inv_A_t2cano: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pose (B,J,3) axis-angle -> A_cano2pose (B,J,4,4) = A_t2pose @ inv_A_t2cano.

    The per-frame `.A` of the reference's body-model call plus sings_hybrid.py:399, computed
    from cached rest joints (they depend only on betas) by one small
    kernel: batch_rodrigues + batch_rigid_transform (body_model/smpl.py:415-513).
    Differentiable w.r.t. pose."""
    squeeze = pose.dim() == 2
    if squeeze:
        pose = pose[None]
    A = _PoseToA.apply(pose, rest_joints, parents, inv_A_t2cano)
    return A[0] if squeeze else A


class _Rot6dConvert(torch.autograd.Function):
    """mode 0: (…,6) -> (…,3,3); mode 1: (…,6) -> (…,3) axis-angle."""

    @staticmethod
    def forward(ctx, d6, mode):
        _need_cuda(d6, "rotation_6d conversion")
        d6_c = _c(d6).reshape(-1, 6)
        n = d6_c.shape[0]
        out = torch.empty((n, 9) if mode == 0 else (n, 3), device=d6.device, dtype=torch.float32)
        st = _lib.raw_stream(d6.device)
        fn = _lib.lib().sgs_rot6d_to_matrix if mode == 0 else _lib.lib().sgs_rot6d_to_axis_angle
        _lib.check(fn(d6_c.data_ptr(), n, out.data_ptr(), st), "sgs_rot6d_to_*")
        ctx.save_for_backward(d6_c)
        ctx.mode, ctx.in_shape = mode, d6.shape
        return out.reshape(d6.shape[:-1] + ((3, 3) if mode == 0 else (3,)))

    @staticmethod
    def backward(ctx, g):
        (d6_c,) = ctx.saved_tensors
        n = d6_c.shape[0]
        g_c = _c(g).reshape(n, -1)
        g6 = torch.empty(n, 6, device=d6_c.device, dtype=torch.float32)
        st = _lib.raw_stream(d6_c.device)
        fn = _lib.lib().sgs_rot6d_to_matrix_bwd if ctx.mode == 0 else _lib.lib().sgs_rot6d_to_axis_angle_bwd
        _lib.check(fn(d6_c.data_ptr(), g_c.data_ptr(), n, g6.data_ptr(), st), "sgs_rot6d_to_*_bwd")
        return g6.reshape(ctx.in_shape), None


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """(…,6) -> (…,3,3), rows b1, b2, b3 of the Gram-Schmidt basis; same name and meaning as
    /root/reference/sings/rec/utils/geometry/rotations.py:545-566.  One kernel each way."""
    return _Rot6dConvert.apply(d6, 0)


def rotation_6d_to_axis_angle(d6: torch.Tensor) -> torch.Tensor:
    """(…,6) -> (…,3) axis-angle: rotations.py:601-603 (6D -> matrix -> quaternion ->
    axis-angle), the per-frame conversion of the stored pose parameters
    (sings_hybrid.py:370-376).  One kernel each way instead of ~40 eager ops."""
    return _Rot6dConvert.apply(d6, 1)


class _DeformGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, xyz, W, rot, scales, smpl_scale, transl, ext_trans, ext_rot, ext_scale,
                want_T, rot6d=False):
        _need_cuda(xyz, "deform_gaussians")
        A_c, xyz_c, W_c, rot_c, sc_c = _c(A), _c(xyz), _c(W), _c(rot), _c(scales)
        ss_c, tr_c = _c(smpl_scale), _c(transl)
        et_c, er_c, es_c = _c(ext_trans), _c(ext_rot), _c(ext_scale)
        B, J = A_c.shape[0], A_c.shape[1]
        N = xyz_c.shape[0]
        dev = xyz_c.device
        xyz_o = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        rotq_o = torch.empty(B, N, 4, device=dev, dtype=torch.float32)
        sc_o = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        T_o = torch.empty(B, N, 4, 4, device=dev, dtype=torch.float32) if want_T else None
        st = _lib.raw_stream(dev)
        fwd = _lib.lib().sgs_lbs_fwd_rot6d if rot6d else _lib.lib().sgs_lbs_fwd
        _lib.check(fwd(
            B, N, J, A_c.data_ptr(), xyz_c.data_ptr(), W_c.data_ptr(), _lib.ptr(rot_c),
            sc_c.data_ptr(), _lib.ptr(ss_c), _lib.ptr(tr_c), _lib.ptr(et_c), _lib.ptr(er_c),
            _lib.ptr(es_c), xyz_o.data_ptr(), rotq_o.data_ptr(), sc_o.data_ptr(), _lib.ptr(T_o), st),
            "sgs_lbs_fwd")
        ctx.save_for_backward(A_c, xyz_c, W_c, rot_c, sc_c, ss_c, tr_c, et_c, er_c, es_c)
        ctx.rot6d = bool(rot6d)
        ctx.shapes = (A.shape, smpl_scale.shape if smpl_scale is not None else None,
                      transl.shape if transl is not None else None)
        if want_T:
            return xyz_o, rotq_o, sc_o, T_o
        return xyz_o, rotq_o, sc_o

    @staticmethod
    def backward(ctx, g_xyz, g_rotq, g_scales, g_T=None):
        A_c, xyz_c, W_c, rot_c, sc_c, ss_c, tr_c, et_c, er_c, es_c = ctx.saved_tensors
        B, J = A_c.shape[0], A_c.shape[1]
        N = xyz_c.shape[0]
        dev = xyz_c.device
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        g_xyz = _c(g_xyz) if g_xyz is not None else z(B, N, 3)
        g_rotq = _c(g_rotq) if g_rotq is not None else z(B, N, 4)
        g_scales = _c(g_scales) if g_scales is not None else z(B, N, 3)
        g_T = _c(g_T)
        d_xyz = torch.empty(N, 3, device=dev, dtype=torch.float32)
        d_rot = None
        if rot_c is not None:
            d_rot = torch.empty((N, 6) if ctx.rot6d else (N, 3, 3), device=dev, dtype=torch.float32)
        d_sc = torch.empty(N, 3, device=dev, dtype=torch.float32)
        d_A = z(B, J, 4, 4)
        d_ss = z(B) if ss_c is not None else None
        d_tr = z(B, 3) if tr_c is not None else None
        st = _lib.raw_stream(dev)
        bwd = _lib.lib().sgs_lbs_bwd_rot6d if ctx.rot6d else _lib.lib().sgs_lbs_bwd
        _lib.check(bwd(
            B, N, J, A_c.data_ptr(), xyz_c.data_ptr(), W_c.data_ptr(), _lib.ptr(rot_c),
            sc_c.data_ptr(), _lib.ptr(ss_c), _lib.ptr(tr_c), _lib.ptr(et_c), _lib.ptr(er_c),
            _lib.ptr(es_c), g_xyz.data_ptr(), g_rotq.data_ptr(), g_scales.data_ptr(),
            _lib.ptr(g_T), d_xyz.data_ptr(), _lib.ptr(d_rot), d_sc.data_ptr(), d_A.data_ptr(), _lib.ptr(d_ss),
            _lib.ptr(d_tr), st), "sgs_lbs_bwd")
        A_shape, ss_shape, tr_shape = ctx.shapes
        return (d_A.reshape(A_shape), d_xyz, None, d_rot, d_sc,
                d_ss.reshape(ss_shape) if d_ss is not None else None,
                d_tr.reshape(tr_shape) if d_tr is not None else None, None, None, None, None, None)


def deform_gaussians(A_cano2pose: torch.Tensor, xyz_canon: torch.Tensor,
                     lbs_weights: torch.Tensor, rotmat_canon: Optional[torch.Tensor],
                     scales: torch.Tensor, smpl_scale: Optional[torch.Tensor] = None,
                     transl: Optional[torch.Tensor] = None,
                     ext_tfs: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                     return_T: bool = False, rot6d_canon: Optional[torch.Tensor] = None):
    """Fused deform of every Gaussian.

    A_cano2pose (J,4,4) or (B,J,4,4); xyz_canon (N,3); lbs_weights (N,J);
    rotmat_canon (N,3,3) or None (identity = the isotropic case, sings_hybrid.py:360) -- or
    pass rot6d_canon (N,6), the stored parameter, and the kernel does rotation_6d_to_matrix
    (sings_hybrid.py:354-356) itself, forward and backward, without the (N,3,3) round trip;
    scales (N,3); smpl_scale (1,) / (B,1) / (B,); transl (3,) / (B,3);
    ext_tfs = (trans (B,3)|(1,3), rotmat (B,3,3)|(1,3,3), scale (B,1)|(1,1)).
    Returns (xyz, rotq, scales[, T]) with a leading B dimension iff A had one."""
    single = A_cano2pose.dim() == 3
    A = A_cano2pose[None] if single else A_cano2pose
    B = A.shape[0]
    ss = None if smpl_scale is None else smpl_scale.reshape(-1).expand(B) if smpl_scale.numel() == 1 else smpl_scale.reshape(B)
    tr = None if transl is None else transl.reshape(-1, 3).expand(B, 3)
    et = er = es = None
    if ext_tfs is not None:
        trans, rotmat, scale = ext_tfs
        et = trans.reshape(-1, 3).expand(B, 3)
        er = rotmat.reshape(-1, 3, 3).expand(B, 3, 3)
        es = scale.reshape(-1).expand(B)
    if rot6d_canon is not None and rotmat_canon is not None:
        raise ValueError("pass rotmat_canon or rot6d_canon, not both")
    if rot6d_canon is not None and rot6d_canon.shape[-1] != 6:
        raise ValueError("rot6d_canon must have shape (N, 6)")
    out = _DeformGaussians.apply(A, xyz_canon, lbs_weights,
                                 rot6d_canon if rot6d_canon is not None else rotmat_canon, scales,
                                 ss, tr, et, er, es, return_T, rot6d_canon is not None)
    if single:
        out = tuple(o[0] for o in out)
    return out


def lbs_extra(A, v_shaped, posedirs, lbs_weights, pose, disable_posedirs: bool = False,
              pose2rot: bool = True):
    """Signature-compatible with the reference's lbs_extra (body_model/lbs.py:16-74) for the
    only configuration SinGS uses (disable_posedirs=True: sings_hybrid.py:57,78,404).
    A (B,J,4,4), v_shaped (B,N,3) [expanded view of one (N,3) tensor], lbs_weights (N,J)
    -> (verts (B,N,3), A, T (B,N,4,4), v_posed, v_shaped)."""
    if not disable_posedirs:
        raise NotImplementedError("pose blend shapes are off in SinGS (disable_posedirs=True); "
                                  "the fused kernel implements that configuration only")
    B, N = v_shaped.shape[0], v_shaped.shape[1]
    if B > 1 and v_shaped.stride(0) != 0:
        raise NotImplementedError("lbs_extra expects v_shaped to be one (N,3) tensor expanded over B")
    xyz_canon = v_shaped[0]
    ident_scales = torch.ones(N, 3, device=v_shaped.device, dtype=torch.float32)
    verts, _, _, T = deform_gaussians(A, xyz_canon, lbs_weights, None, ident_scales, return_T=True)
    return verts, A, T, v_shaped, v_shaped

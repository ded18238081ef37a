"""CPU (torch) restatement of the reference's per-frame avatar deformer.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by sings_b200/ or diff_gaussian_rasterization/.

Parity status: PINNED.  tests/golden/lbs_golden_*.npz hold inputs and outputs (values and
autograd gradients) produced by importing the reference's OWN code in the build container
(tests/golden/make_lbs_golden.py: lbs_extra from /root/reference/sings/rec/utils/body_model/
lbs.py:16-74, matrix_to_quaternion / quaternion_multiply from .../geometry/rotations.py:98-149,
393-407, batch_rodrigues / batch_rigid_transform from .../body_model/smpl.py:415-513;
tests/golden/make_rot6d_golden.py: rotation_6d_to_matrix / rotation_6d_to_axis_angle from
.../geometry/rotations.py:545-566, 601-603 -> tests/golden/rot6d_golden_*.npz);
tests/test_oracle_lbs.py checks every function below against those vectors.

Every function works on torch tensors of any float dtype and is autograd-differentiable,
so float64 gradients of the restatement are the gradient truth for the CUDA backward.
"""
from __future__ import annotations

import torch


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """(B,3) axis-angle -> (B,3,3).  smpl.py:415-446: theta = ||r + 1e-8||, K = skew(r/theta),
    R = I + sin(theta) K + (1 - cos(theta)) K^2."""
    angle = torch.linalg.vector_norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).reshape(-1, 3, 3)
    s = torch.sin(angle)[:, :, None]
    c = torch.cos(angle)[:, :, None]
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return eye + s * K + (1 - c) * (K @ K)


def batch_rigid_transform(rot_mats: torch.Tensor, joints: torch.Tensor, parents) -> torch.Tensor:
    """(B,J,3,3), (B,J,3) -> relative transforms A (B,J,4,4).  smpl.py:462-513:
    G_0 = [R_0 | j_0], G_j = G_parent [R_j | j_j - j_parent], A_j = G_j - pad(G_j [j_j; 0])."""
    B, J = joints.shape[:2]
    parents = [int(p) for p in parents]
    rel = joints.clone()
    rel[:, 1:] = joints[:, 1:] - joints[:, parents[1:]]
    bottom = torch.zeros(B, J, 1, 4, dtype=joints.dtype, device=joints.device)
    bottom[..., 3] = 1
    local = torch.cat([torch.cat([rot_mats, rel[..., None]], dim=-1), bottom], dim=-2)
    chain = [local[:, 0]]
    for j in range(1, J):
        chain.append(chain[parents[j]] @ local[:, j])
    G = torch.stack(chain, dim=1)
    jh = torch.cat([joints, torch.zeros_like(joints[..., :1])], dim=-1)[..., None]
    corr = G @ jh                                           # (B,J,4,1)
    pad = torch.zeros_like(G)
    pad[..., 3:4] = corr
    return G - pad


def pose_to_A(pose: torch.Tensor, rest_joints: torch.Tensor, parents,
              inv_A_t2cano: torch.Tensor | None = None) -> torch.Tensor:
    """pose (B,J,3) axis-angle -> A_cano2pose (B,J,4,4): the `.A` of the body-model call
    (sings_hybrid.py:390-398; lbs.py:126-171 with rest joints cached)
    followed by `A_t2pose @ inv_A_t2cano` (sings_hybrid.py:399, :525)."""
    B, J = pose.shape[:2]
    Rm = batch_rodrigues(pose.reshape(-1, 3)).reshape(B, J, 3, 3)
    A = batch_rigid_transform(Rm, rest_joints[None].expand(B, -1, -1), parents)
    if inv_A_t2cano is not None:
        A = A @ inv_A_t2cano[None]
    return A


def lbs_extra(A: torch.Tensor, v_shaped: torch.Tensor, lbs_weights: torch.Tensor):
    """lbs.py:16-74 with disable_posedirs=True (hard-wired at sings_hybrid.py:57,78,404):
    T = W @ A.view(J,16); verts = (T [v;1])[:3].  Returns (verts (B,N,3), T (B,N,4,4))."""
    B, J = A.shape[:2]
    T = (lbs_weights[None].expand(B, -1, -1) @ A.reshape(B, J, 16)).reshape(B, -1, 4, 4)
    vh = torch.cat([v_shaped, torch.ones_like(v_shaped[..., :1])], dim=-1)
    verts = (T @ vh[..., None])[..., :3, 0]
    return verts, T


def matrix_to_quaternion(m: torch.Tensor) -> torch.Tensor:
    """(...,3,3) -> (...,4) real part first; rotations.py:98-149: four candidates, divided by
    2 max(q_abs, 0.1), pick argmax(q_abs).  sqrt has a zero sub-gradient at 0
    (rotations.py:87-95).  The result is NOT normalised."""
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    arg = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22,
                       1 - m00 - m11 + m22], dim=-1)
    pos = arg > 0
    q_abs = torch.where(pos, torch.sqrt(torch.where(pos, arg, torch.ones_like(arg))),
                        torch.zeros_like(arg))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4)).squeeze(-2)


def quaternion_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product, then sign-standardised to real part >= 0 (rotations.py:357-407)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    q = torch.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], dim=-1)
    return torch.where(q[..., 0:1] < 0, -q, q)


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    """(...,6) -> (...,3,3), rows b1, b2, b3 (rotations.py:545-566): Gram-Schmidt with
    F.normalize semantics (divide by max(norm, 1e-12))."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = a1 / a1.norm(dim=-1, keepdim=True).clamp(min=1e-12)
    u = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = u / u.norm(dim=-1, keepdim=True).clamp(min=1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def quaternion_to_axis_angle(q: torch.Tensor) -> torch.Tensor:
    """(...,4) real first -> (...,3) (rotations.py:514-542): half = atan2(|v|, w), angle = 2 half,
    v / (sin(half)/angle), with the series 0.5 - angle^2/48 where |angle| < 1e-6."""
    norms = q[..., 1:].norm(dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    angle = 2 * half
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    s = torch.where(small, 0.5 - (angle * angle) / 48, torch.sin(torch.where(small, torch.ones_like(half), half)) / safe)
    return q[..., 1:] / s


def rotation_6d_to_axis_angle(d6: torch.Tensor) -> torch.Tensor:
    """rotations.py:601-603: 6D -> matrix -> quaternion (:98-149) -> axis-angle; the per-frame
    conversion of the stored pose parameters (sings_hybrid.py:370-376)."""
    return quaternion_to_axis_angle(matrix_to_quaternion(rotation_6d_to_matrix(d6)))


def deform(A_cano2pose, xyz_canon, lbs_weights, scales, rotmat_canon=None, smpl_scale=None,
           transl=None, ext_tfs=None, rot6d_canon=None):
    """The deform segment of SinGS.forward / forward_chunk (sings_hybrid.py:398-428, :525-552;
    SURVEY.md Appendix B steps 2-6), batched over B frames.

    A_cano2pose (B,J,4,4); xyz_canon (N,3); lbs_weights (N,J); scales (N,3);
    rotmat_canon (N,3,3) or None (= identity, the isotropic case sings_hybrid.py:360);
    smpl_scale (B,1) or None; transl (B,3) or None;
    ext_tfs = (trans (B,3), rotmat (B,3,3), scale (B,1)) or None.
    Returns xyz (B,N,3), rotq (B,N,4), scales (B,N,3), T (B,N,4,4).
    """
    B = A_cano2pose.shape[0]
    N = xyz_canon.shape[0]
    if rot6d_canon is not None:      # sings_hybrid.py:354-356
        rotmat_canon = rotation_6d_to_matrix(rot6d_canon)
    xyz, T = lbs_extra(A_cano2pose, xyz_canon[None].expand(B, -1, -1), lbs_weights)
    sc = scales[None].expand(B, -1, -1)
    if smpl_scale is not None:
        xyz = xyz * smpl_scale[:, None, :]
        sc = sc * smpl_scale[:, None, :]
    if transl is not None:
        xyz = xyz + transl[:, None, :]
    if rotmat_canon is None:
        rotmat_canon = torch.eye(3, dtype=xyz.dtype, device=xyz.device)[None].expand(N, -1, -1)
    Rdef = T[..., :3, :3] @ rotmat_canon[None]
    q = matrix_to_quaternion(Rdef)
    if ext_tfs is not None:
        trans, rotmat, scale = ext_tfs
        xyz = trans[:, None, :] + scale[:, None, :] * (rotmat[:, None] @ xyz[..., None]).squeeze(-1)
        sc = scale[:, None, :] * sc
        q = quaternion_multiply(matrix_to_quaternion(rotmat)[:, None, :], q)
    return xyz, q, sc, T

"""GPU parity tests of the fused deformer (sings_b200.deform) against the reference-generated
golden vectors and the CPU oracle.  LBS is floating point with a different summation order
than the reference's cuBLAS/torch path, so parity is tolerance-based: 1e-5 absolute on
positions (metres) / quaternions, 1e-3 relative on gradients."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import assert_grad_close, rel_err
from oracle import lbs_oracle as lo
from sings_b200 import deform
from sings_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lbs_golden_*_f32.npz")))
VAL_TOL, GRAD_TOL = 1e-5, 1e-3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[11:-8] for p in GOLD])
def test_against_reference_golden(path):
    g = {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    g64 = {k: torch.from_numpy(v) for k, v in np.load(path.replace("_f32", "_f64")).items()}
    dev = "cuda"
    c = lambda k, rg=False: g[k].to(dev).requires_grad_(rg) if k in g else None
    iso = bool(g["isotropic"])
    pose = c("pose", True)
    A = deform.pose_to_A(pose, c("rest"), g["parents"], c("inv_A_t2cano"))
    assert (A.detach().cpu() - g["A_cano2pose"]).abs().max() < VAL_TOL
    ext = (c("ext_trans"), c("ext_rotmat"), c("ext_scale")) if "ext_trans" in g else None
    xyz_c, sc_c = c("xyz_canon", True), c("scales", True)
    rc = None if iso else c("rotmat_canon", True)
    ss, tr = c("smpl_scale", True), c("transl", True)
    xyz, q, sco, T = deform.deform_gaussians(A, xyz_c, c("lbs_weights"), rc, sc_c, ss, tr, ext, return_T=True)
    assert (xyz.detach().cpu() - g["xyz"]).abs().max() < VAL_TOL
    assert (q.detach().cpu() - g["rotq"]).abs().max() < VAL_TOL
    assert (sco.detach().cpu() - g["scales_out"]).abs().max() < VAL_TOL
    assert (T.detach().cpu() - g["T"]).abs().max() < VAL_TOL
    loss = (xyz * c("gx")).sum() + (q * c("gq")).sum() + (sco * c("gs")).sum()
    loss.backward()
    ref = g64     # float64 autograd of the reference code = gradient truth
    checks = [("d_pose", pose), ("d_xyz_canon", xyz_c), ("d_scales", sc_c), ("d_smpl_scale", ss), ("d_transl", tr)]
    if rc is not None:
        checks.append(("d_rotmat_canon", rc))
    for name, t in checks:
        assert_grad_close(t.grad.cpu().numpy(), ref[name].numpy(), name, tol=GRAD_TOL)


def test_lbs_extra_signature_and_T_gradient():
    """lbs_extra keeps the reference signature (lbs.py:16-24, 74) and its T output carries
    gradient to A, as sings_hybrid.py:418 needs."""
    g = {k: torch.from_numpy(v) for k, v in np.load(GOLD[0]).items()}
    dev = "cuda"
    A = g["A_cano2pose"].to(dev).requires_grad_(True)
    v = g["xyz_canon"].to(dev)[None].expand(A.shape[0], -1, -1)
    verts, A_out, T, v_posed, v_shaped = deform.lbs_extra(A, v, None, g["lbs_weights"].to(dev), None,
                                                          disable_posedirs=True, pose2rot=True)
    A_cpu = g["A_cano2pose"].clone().requires_grad_(True)
    verts_o, T_o = lo.lbs_extra(A_cpu, g["xyz_canon"][None].expand(A.shape[0], -1, -1), g["lbs_weights"])
    assert (verts.detach().cpu() - verts_o.detach()).abs().max() < VAL_TOL
    assert (T.detach().cpu() - T_o.detach()).abs().max() < VAL_TOL
    gT = torch.randn(T_o.shape, generator=torch.Generator().manual_seed(0))
    gv = torch.randn(verts_o.shape, generator=torch.Generator().manual_seed(1))
    ((T * gT.to(dev)).sum() + (verts * gv.to(dev)).sum()).backward()
    ((T_o * gT).sum() + (verts_o * gv).sum()).backward()
    assert rel_err(A.grad.cpu().numpy()[:, :, :3], A_cpu.grad.numpy()[:, :, :3]) < GRAD_TOL
    with pytest.raises(NotImplementedError):
        deform.lbs_extra(A, v, None, g["lbs_weights"].to(dev), None, disable_posedirs=False)


@pytest.mark.parametrize("N,J,B,iso", [(1, 24, 1, False), (257, 24, 1, True), (5003, 52, 2, False),
                                       (200_000, 24, 1, False), (4099, 24, 16, True)])
def test_ragged_sizes_vs_oracle(N, J, B, iso):
    """N not a multiple of the 256-row tile (TMA bulk size changes per CTA), chunked B=16 frames
    (forward_chunk, sings_hybrid.py:474-569), and the full 200k config."""
    av = syn.make_avatar(N, J, seed=N % 97, isotropic=iso)
    pose = torch.stack([torch.from_numpy(syn.random_pose(J, seed=b)) for b in range(B)])
    t = torch.from_numpy
    A = lo.pose_to_A(pose, t(av.rest), av.parents, t(av.inv_A_t2cano))
    transl = torch.randn(B, 3, generator=torch.Generator().manual_seed(0))
    rot = None if iso else t(av.rotmat_canon)
    xo, qo, so, _ = lo.deform(A, t(av.xyz_canon), t(av.lbs_weights), t(av.scales), rot, None, transl)
    d = lambda a: None if a is None else a.cuda()
    x, q, s = deform.deform_gaussians(d(A), d(t(av.xyz_canon)), d(t(av.lbs_weights)), d(rot), d(t(av.scales)),
                                      None, d(transl))
    assert (x.cpu() - xo).abs().max() < 2e-5           # metres, at |x| ~ 10
    assert (s.cpu() - so).abs().max() < 1e-6
    # the candidate argmax may flip between near-tied candidates: compare up to the induced rotation
    from oracle.raster_ref64 import quat_to_R
    Rm = quat_to_R(q.cpu().reshape(-1, 4).double())
    Ro = quat_to_R(qo.reshape(-1, 4).double())
    assert (Rm - Ro).abs().max() < 1e-4


def test_pose_to_A_identity_and_batch():
    rest, parents, _ = syn.skeleton(24)
    A = deform.pose_to_A(torch.zeros(3, 24, 3, device="cuda"), torch.tensor(rest, dtype=torch.float32, device="cuda"),
                         torch.from_numpy(parents), None)
    assert (A.cpu() - torch.eye(4).expand(3, 24, 4, 4)).abs().max() < 1e-6


def test_rot6d_conversions_against_reference_golden():
    """rotation_6d_to_matrix / rotation_6d_to_axis_angle kernels (rotations.py:545-566, 601-603)
    against vectors made by the reference's own functions; gradients against its float64 autograd."""
    base = os.path.join(os.path.dirname(__file__), "golden", "rot6d_golden_")
    g = {k: torch.from_numpy(v) for k, v in np.load(base + "f32.npz").items()}
    g64 = {k: torch.from_numpy(v) for k, v in np.load(base + "f64.npz").items()}
    d6 = g["d6"].cuda().requires_grad_(True)
    R = deform.rotation_6d_to_matrix(d6)
    aa = deform.rotation_6d_to_axis_angle(d6)
    assert R.shape == (d6.shape[0], 3, 3) and aa.shape == (d6.shape[0], 3)
    assert (R.detach().cpu() - g["R"]).abs().max() < VAL_TOL
    # near pi the axis-angle is ill-conditioned in float32: compare with the float32 reference run
    assert (aa.detach().cpu() - g["aa"]).abs().max() < 2e-4
    well = (g64["aa"].norm(dim=-1) < 2.8)
    assert (aa.detach().cpu()[well] - g64["aa"][well].float()).abs().max() < 2e-5
    (dR,) = torch.autograd.grad((R * g["gR"].cuda()).sum(), d6, retain_graph=True)
    (daa,) = torch.autograd.grad((aa * g["gaa"].cuda()).sum(), d6)
    assert rel_err(dR.cpu().numpy(), g64["d_d6_from_R"].numpy()) < GRAD_TOL
    assert rel_err(daa.cpu().numpy()[well], g64["d_d6_from_aa"].numpy()[well]) < GRAD_TOL
    # batched leading dimensions, as the model stores them: (frames, 23, 6)
    d = torch.randn(4, 23, 6, device="cuda")
    assert deform.rotation_6d_to_axis_angle(d).shape == (4, 23, 3)
    assert torch.equal(deform.rotation_6d_to_matrix(d).reshape(-1, 3, 3), deform.rotation_6d_to_matrix(d.reshape(-1, 6)))


@pytest.mark.parametrize("N,J,B,ext", [(301, 24, 3, True), (5003, 52, 2, False), (200_000, 24, 1, False)])
def test_deform_from_rot6d_vs_oracle(N, J, B, ext):
    """The stored 6D canonical rotation goes straight into the LBS kernels (sings_hybrid.py:354-356
    fused away): values equal the matrix path bit for bit, gradients w.r.t. the 6D parameter
    match float64 autograd of the oracle."""
    av = syn.make_avatar(N, J, seed=7)
    gen = torch.Generator().manual_seed(N)
    t = torch.from_numpy
    d6 = torch.randn(N, 6, generator=gen)
    pose = torch.stack([t(syn.random_pose(J, seed=b)) for b in range(B)])
    A = lo.pose_to_A(pose, t(av.rest), av.parents, t(av.inv_A_t2cano))
    transl = torch.randn(B, 3, generator=gen)
    ss = 1.0 + 0.1 * torch.rand(B, 1, generator=gen)
    ext_tfs = None
    if ext:
        ext_tfs = (torch.randn(B, 3, generator=gen), lo.batch_rodrigues(torch.randn(B, 3, generator=gen)),
                   0.5 + torch.rand(B, 1, generator=gen))
    gx, gq, gs = (torch.randn(B, N, k, generator=gen) for k in (3, 4, 3))
    # float64 oracle
    dd = lambda a: a.double()
    d6o, xo_c, so_c = dd(d6).requires_grad_(True), dd(t(av.xyz_canon)).requires_grad_(True), dd(t(av.scales)).requires_grad_(True)
    xo, qo, sco, _ = lo.deform(dd(A), xo_c, dd(t(av.lbs_weights)), so_c, None, dd(ss), dd(transl),
                               tuple(dd(e) for e in ext_tfs) if ext else None, rot6d_canon=d6o)
    ((xo * gx).sum() + (qo * gq).sum() + (sco * gs).sum()).backward()
    # CUDA, 6D in
    c = lambda a: a.cuda()
    d6c, x_c, s_c = c(d6).requires_grad_(True), c(t(av.xyz_canon)).requires_grad_(True), c(t(av.scales)).requires_grad_(True)
    ext_c = tuple(c(e) for e in ext_tfs) if ext else None
    x, q, s = deform.deform_gaussians(c(A), x_c, c(t(av.lbs_weights)), None, s_c, c(ss), c(transl), ext_c,
                                      rot6d_canon=d6c)
    ((x * c(gx)).sum() + (q * c(gq)).sum() + (s * c(gs)).sum()).backward()
    # CUDA, matrix in (the kernel pair, composed)
    with torch.no_grad():
        Rm = deform.rotation_6d_to_matrix(c(d6))
        x2, q2, s2 = deform.deform_gaussians(c(A), c(t(av.xyz_canon)), c(t(av.lbs_weights)), Rm, c(t(av.scales)),
                                             c(ss), c(transl), ext_c)
    assert torch.equal(x.detach(), x2) and torch.equal(q.detach(), q2) and torch.equal(s.detach(), s2)
    assert (x.detach().cpu() - xo.detach().float()).abs().max() < 2e-5
    from oracle.raster_ref64 import quat_to_R
    assert (quat_to_R(q.detach().cpu().reshape(-1, 4).double()) - quat_to_R(qo.detach().reshape(-1, 4))).abs().max() < 1e-4
    assert d6c.grad.shape == (N, 6)
    for name, got, ref in (("d_rot6d", d6c.grad, d6o.grad), ("d_xyz", x_c.grad, xo_c.grad), ("d_scales", s_c.grad, so_c.grad)):
        e = rel_err(got.cpu().numpy(), ref.numpy())
        assert e < GRAD_TOL, f"{name}: {e}"
    with pytest.raises(ValueError):
        deform.deform_gaussians(c(A), x_c, c(t(av.lbs_weights)), Rm, s_c, rot6d_canon=d6c)

"""CPU: the deformer oracle (oracle/lbs_oracle.py) against golden vectors produced by running
the reference's own code (tests/golden/make_lbs_golden.py; lbs.py:16-74, rotations.py:98-149,
393-407, smpl.py:415-513 composed as sings_hybrid.py:525-552)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import lbs_oracle as lo

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lbs_golden_*.npz")))


def load(path):
    return {k: torch.from_numpy(v) for k, v in np.load(path).items()}


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[11:-4] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = load(path)
    f64 = path.endswith("f64.npz")
    tol = 1e-12 if f64 else 2e-6
    ext = (g["ext_trans"], g["ext_rotmat"], g["ext_scale"]) if "ext_trans" in g else None
    leaves = [g["pose"], g["xyz_canon"], g["scales"], g["rotmat_canon"], g["smpl_scale"], g["transl"]]
    for t in leaves:
        t.requires_grad_(True)
    A = lo.pose_to_A(g["pose"], g["rest"], g["parents"], g["inv_A_t2cano"])
    xyz, q, sc, T = lo.deform(A, g["xyz_canon"], g["lbs_weights"], g["scales"], g["rotmat_canon"],
                              g["smpl_scale"], g["transl"], ext)
    assert (A - g["A_cano2pose"]).abs().max() <= tol
    assert (xyz - g["xyz"]).abs().max() <= tol * 10
    assert (q - g["rotq"]).abs().max() <= tol * 10
    assert (sc - g["scales_out"]).abs().max() <= tol
    assert (T - g["T"]).abs().max() <= tol * 10
    loss = (xyz * g["gx"]).sum() + (q * g["gq"]).sum() + (sc * g["gs"]).sum()
    grads = torch.autograd.grad(loss, leaves)
    names = ["d_pose", "d_xyz_canon", "d_scales", "d_rotmat_canon", "d_smpl_scale", "d_transl"]
    for gr, n in zip(grads, names):
        ref = g[n]
        assert (gr - ref).abs().max() <= (1e-10 if f64 else 1e-3) * (ref.abs().max() + 1e-12), n


def test_quaternion_not_normalised_and_identity_pose():
    """Blended rotations are not orthonormal, so the quaternion is not unit;
    zero pose with identity inv_A gives A = I and LBS is the identity map."""
    g = load(GOLD[0])
    n = g["rotq"].norm(dim=-1)
    assert n.min() < 0.999 and n.max() <= 1.0 + 1e-5
    J = g["rest"].shape[0]
    A = lo.pose_to_A(torch.zeros(1, J, 3, dtype=g["rest"].dtype), g["rest"], g["parents"])
    eye = torch.eye(4, dtype=A.dtype).expand(1, J, 4, 4)
    assert (A - eye).abs().max() < 1e-6
    xyz, q, sc, _ = lo.deform(A, g["xyz_canon"], g["lbs_weights"], g["scales"])
    assert (xyz[0] - g["xyz_canon"]).abs().max() < 1e-5
    assert (q[0] - torch.tensor([1.0, 0, 0, 0], dtype=q.dtype)).abs().max() < 1e-5


ROT6D = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "rot6d_golden_*.npz")))


@pytest.mark.parametrize("path", ROT6D, ids=[os.path.basename(p)[13:-4] for p in ROT6D])
def test_rot6d_oracle_matches_reference_golden(path):
    """rotation_6d_to_matrix / rotation_6d_to_axis_angle (rotations.py:545-566, 601-603) and their
    autograd gradients, incl. far-from-unit inputs, small angles, the identity and angles near pi."""
    g = load(path)
    f64 = path.endswith("f64.npz")
    tol = 1e-12 if f64 else 2e-6
    d6 = g["d6"].requires_grad_(True)
    R = lo.rotation_6d_to_matrix(d6)
    aa = lo.rotation_6d_to_axis_angle(d6)
    assert (R - g["R"]).abs().max() <= tol
    assert (aa - g["aa"]).abs().max() <= tol * 10
    d_R = torch.autograd.grad((R * g["gR"]).sum(), d6, retain_graph=True)[0]
    d_aa = torch.autograd.grad((aa * g["gaa"]).sum(), d6)[0]
    for got, name in ((d_R, "d_d6_from_R"), (d_aa, "d_d6_from_aa")):
        ref = g[name]
        assert torch.isfinite(got).all()
        assert (got - ref).abs().max() <= (1e-9 if f64 else 2e-3) * (ref.abs().max() + 1e-12), name


def test_deform_rot6d_equals_matrix_path():
    """deform(rot6d_canon=d6) == deform(rotmat_canon=rotation_6d_to_matrix(d6)) (sings_hybrid.py:354-356)."""
    g = load([p for p in GOLD if "j24_aniso_ext_f64" in p][0])
    N = g["xyz_canon"].shape[0]
    d6 = torch.randn(N, 6, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    A = g["A_cano2pose"]
    a = lo.deform(A, g["xyz_canon"], g["lbs_weights"], g["scales"], rot6d_canon=d6)
    b = lo.deform(A, g["xyz_canon"], g["lbs_weights"], g["scales"], rotmat_canon=lo.rotation_6d_to_matrix(d6))
    for x, y in zip(a, b):
        assert torch.equal(x, y)

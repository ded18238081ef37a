"""CPU: the oracle of the non-image loss terms (oracle/reg_oracle.py) against golden vectors produced by
the reference's own classes (tests/golden/make_reg_golden.py), and the host-side construction of the
sparse operator (sings_b200/regularizers.py -- index arithmetic, no kernels) against both."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import reg_oracle as ro
from sings_b200 import regularizers as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
POSITION_W = {'head-neck': 0.5, 'spine': 0.75, 'leftUpArm': 1., 'rightUpArm': 1., 'leftDownArm': 1., 'rightDownArm': 1.,
              'leftHand': 1.5, 'rightHand': 1.5, 'hips': 1., 'leftUpLeg': 1., 'rightUpLeg': 1., 'leftDownLeg': 1.,
              'rightDownLeg': 1., 'leftFoot': 0.75, 'rightFoot': 0.75}
COLOR_W = {k: (1.0 if k in ('leftDownArm', 'rightDownArm', 'leftHand', 'rightHand') else 0.0) for k in POSITION_W}


def load(name):
    return np.load(os.path.join(GOLD, f"reg_golden_{name}.npz"))


def test_goldens_are_committed():
    assert len(glob.glob(os.path.join(GOLD, "reg_golden_*.npz"))) == 7


def test_parse_weights_matches_reference_table():
    z = load("region_a")
    np.testing.assert_array_equal(R.parse_weights(POSITION_W), z["weights_pos"])
    np.testing.assert_array_equal(R.parse_weights(COLOR_W), z["weights_col"])
    np.testing.assert_array_equal(ro.parse_weights(POSITION_W, R.REGION_LABEL_MAP), z["weights_pos"])
    np.testing.assert_array_equal(R.parse_weights(None), np.ones(15))
    np.testing.assert_array_equal(R.parse_weights([1, 2, 3]), np.array([1.0, 2.0, 3.0]))


@pytest.mark.parametrize("name", ["region_a", "region_b"])
@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-6)])
def test_region_oracle_vs_reference(name, tag, dt, tol):
    z = load(name)
    verts, edges, labels = torch.from_numpy(z["verts"]), torch.from_numpy(z["edges"]), torch.from_numpy(z["labels"])
    for wname, w, key in (("pos", POSITION_W, "xyz"), ("col", COLOR_W, "shs")):
        lap = ro.RegionLaplacian(verts.to(dt), edges, labels, ro.parse_weights(w, R.REGION_LABEL_MAP))
        src = torch.from_numpy(z[key]).to(dt).requires_grad_(True)
        x = src if key == "xyz" else src[:, 0]
        loss = lap.forward(x)
        g, = torch.autograd.grad(loss, src)
        assert abs(float(loss) - float(z[f"loss_{wname}_{tag}"])) <= tol * abs(float(z[f"loss_{wname}_{tag}"]))
        ref = z[f"grad_{wname}_{tag}"]
        assert np.abs(g.numpy() - ref).max() <= tol * np.abs(ref).max() + 1e-30
        if wname == "pos":
            src2 = torch.from_numpy(z["xyz"]).to(dt).requires_grad_(True)
            lh = lap.forward_hands(src2)
            gh, = torch.autograd.grad(lh, src2)
            assert abs(float(lh) - float(z[f"loss_hand_{tag}"])) <= tol * abs(float(z[f"loss_hand_{tag}"]))
            assert np.abs(gh.numpy() - z[f"grad_hand_{tag}"]).max() <= tol * np.abs(z[f"grad_hand_{tag}"]).max()


@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-6)])
def test_pcd_and_l2norm_oracle_vs_reference(tag, dt, tol):
    z = load("pcd")
    x = torch.from_numpy(z["pts"]).to(dt).requires_grad_(True)
    loss = ro.pcd_laplacian_smoothing(x, torch.from_numpy(z["edges"]))
    g, = torch.autograd.grad(loss, x)
    assert abs(float(loss) - float(z[f"loss_{tag}"])) <= tol * float(z[f"loss_{tag}"])
    assert np.abs(g.numpy() - z[f"grad_{tag}"]).max() <= tol * np.abs(z[f"grad_{tag}"]).max()
    for name in ("l2_a", "l2_b", "l2_c"):
        z = load(name)
        c = z["cfg"]
        kw = dict(lambda_xyz_offsets=c[0], lambda_scales_diff=c[1], lambda_max_scale=c[2], max_scale_threshold=c[3],
                  lambda_min_opacity=c[4], min_opacity_threshold=c[5])
        o = torch.from_numpy(z["xyz_offsets"]).to(dt).requires_grad_(True)
        s = torch.from_numpy(z["scales"]).to(dt).requires_grad_(True)
        p = torch.from_numpy(z["opacity"]).to(dt).requires_grad_(True)
        d = {"xyz_offsets": o, "scales": s}
        if bool(z["has_opacity"]):
            d["opacity"] = p
        loss = ro.l2norm(d, **kw)
        go, gs = torch.autograd.grad(loss, [o, s])
        assert abs(float(loss) - float(z[f"loss_{tag}"])) <= tol * float(z[f"loss_{tag}"])
        assert np.abs(go.numpy() - z[f"grad_off_{tag}"]).max() <= tol * np.abs(z[f"grad_off_{tag}"]).max()
        assert np.abs(gs.numpy() - z[f"grad_scales_{tag}"]).max() <= tol * np.abs(z[f"grad_scales_{tag}"]).max()


@pytest.mark.parametrize("name", ["region_a", "region_b"])
def test_region_operator_matches_the_reference_operators(name):
    """One CSR operator over all vertices == the reference's per-region matrices placed where forward() applies
    them -- including case b, where two vertices have no edge inside their region and the reference's
    renumbering shifts the rows of those regions."""
    z = load(name)
    edges, labels = torch.from_numpy(z["edges"]), torch.from_numpy(z["labels"])
    op, n_region = R.region_laplacian(labels, edges)
    np.testing.assert_array_equal(n_region.numpy(), np.bincount(z["labels"], minlength=15))
    D = op.to_dense().numpy()
    np.testing.assert_allclose(D, z["L_dense"], rtol=0, atol=1e-7)
    if name == "region_b":
        # the quirk is really exercised: some region's used-vertex numbering differs from its membership numbering
        used = np.zeros(len(z["labels"]), bool)
        el = z["labels"][z["edges"]]
        used[z["edges"][el[:, 0] == el[:, 1]].reshape(-1)] = True
        assert (~used).sum() >= 2
    # the transposed CSR is the transpose
    Dt = torch.zeros_like(op.to_dense())
    counts = (op.t_ptr[1:] - op.t_ptr[:-1]).long()
    cols = torch.repeat_interleave(torch.arange(op.n), counts)
    Dt[op.t_row.long(), cols] = op.t_val
    np.testing.assert_array_equal(Dt.numpy(), D)
    assert op.row_ptr.dtype == torch.int32 and op.col_idx.dtype == torch.int32 and op.vals.dtype == torch.float32
    assert int(op.row_ptr[-1]) == op.nnz == int(op.t_ptr[-1])
    # rows are sorted by column (coalesced), no duplicates
    for r in (0, op.n // 2, op.n - 1):
        c = op.col_idx[int(op.row_ptr[r]):int(op.row_ptr[r + 1])].numpy()
        assert (np.diff(c) > 0).all()


def test_plain_laplacian_operator_matches_pytorch3d_definition():
    z = load("pcd")
    pts, edges = torch.from_numpy(z["pts"]), torch.from_numpy(z["edges"])
    op = R.laplacian(pts, edges)
    np.testing.assert_allclose(op.to_dense().numpy(), ro.laplacian(pts, edges).numpy(), rtol=0, atol=1e-7)
    # repeated and self edges add up like a COO tensor
    e = torch.tensor([[0, 1], [1, 0], [0, 1], [2, 2], [3, 0]])
    v = torch.zeros(5, 3)
    np.testing.assert_allclose(R.laplacian(v, e).to_dense().numpy(), ro.laplacian(v, e).numpy(), rtol=0, atol=1e-7)
    assert float(R.laplacian(v, e).to_dense()[4, 4]) == -1.0          # isolated vertex: only the diagonal


def test_sparse_oracle_variant_equals_the_dense_one():
    z = load("region_b")
    verts, edges, labels = torch.from_numpy(z["verts"]), torch.from_numpy(z["edges"]), torch.from_numpy(z["labels"])
    x = torch.from_numpy(z["xyz"])
    a = ro.RegionLaplacian(verts, edges, labels, z["weights_pos"], sparse=True)
    b = ro.RegionLaplacian(verts, edges, labels, z["weights_pos"])
    assert abs(float(a.forward(x)) - float(b.forward(x))) <= 1e-6 * float(b.forward(x))
    assert abs(float(a.forward(x)) - float(z["loss_pos_f32"])) <= 1e-6 * float(z["loss_pos_f32"])
    for La, Lb in zip(a.laplacians, b.laplacians):
        np.testing.assert_allclose(La.to_dense().numpy(), Lb.numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-9)])
def test_scale_edge_loss_oracle_vs_reference(tag, dt, tol):
    """oracle/knn_oracle.py::gaussians_edge_loss against the reference's GaussiansEdgeLoss (run with a brute-force
    knn_points): loss and its gradient to the scales."""
    from oracle import knn_oracle as ko
    z = load("edge")
    x = torch.from_numpy(z["xyz_canon"]).to(dt)
    s = torch.from_numpy(z["scales"]).to(dt).requires_grad_(True)
    loss = ko.gaussians_edge_loss({"xyz_canon": x, "scales": s}, K=9)
    g, = torch.autograd.grad(loss, s)
    assert abs(float(loss) - float(z[f"loss_{tag}"])) <= tol * float(z[f"loss_{tag}"])
    assert np.abs(g.numpy() - z[f"grad_scales_{tag}"]).max() <= tol * np.abs(z[f"grad_scales_{tag}"]).max()


def test_label_checks_and_no_cpu_path():
    edges = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(R.SgsError):
        R.region_laplacian(torch.tensor([0, 2, 2]), edges)             # label 1 missing
    with pytest.raises(R.SgsError):
        R.region_laplacian(torch.tensor([-1, 0, 0]), edges)
    with pytest.raises(NotImplementedError):
        R.RegionLaplacianLoss_v2(torch.zeros(3, 3), edges, torch.tensor([0, 0, 0]), laplacian_type="cotangent")
    lap = R.RegionLaplacianLoss_v2(torch.zeros(3, 3), edges, torch.tensor([0, 0, 0]), region_weights=[2.0])
    with pytest.raises(R.SgsError):
        lap(torch.zeros(3, 3))                                         # CPU tensor: no fallback
    with pytest.raises(R.SgsError):
        R.L2Norm()({"xyz_offsets": torch.zeros(4, 3), "scales": torch.ones(4, 3)})
    with pytest.raises(R.SgsError):
        R.pcd_laplacian_smoothing(torch.zeros(3, 3), edges)
    src = open(R.__file__).read()
    assert "import oracle" not in src and "from oracle" not in src

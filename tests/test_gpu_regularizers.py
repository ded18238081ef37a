"""Region Laplacians, point-cloud Laplacian smoothing and L2Norm (sings_b200/csrc/regularizers.cu through
the C ABI) against the golden vectors produced by the reference's own classes and against the CPU oracle.
Floating point: loss within 1e-5 relative of the reference's float64 value, gradients within 1e-4 of
their largest element and 99.9 % of the elements within 1e-3 relative (tests/helpers.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_grad_close

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOSS_RTOL, GRAD_TOL = 1e-5, 1e-4
POSITION_W = {'head-neck': 0.5, 'spine': 0.75, 'leftUpArm': 1., 'rightUpArm': 1., 'leftDownArm': 1., 'rightDownArm': 1.,
              'leftHand': 1.5, 'rightHand': 1.5, 'hips': 1., 'leftUpLeg': 1., 'rightUpLeg': 1., 'leftDownLeg': 1.,
              'rightDownLeg': 1., 'leftFoot': 0.75, 'rightFoot': 0.75}
COLOR_W = {k: (1.0 if k in ('leftDownArm', 'rightDownArm', 'leftHand', 'rightHand') else 0.0) for k in POSITION_W}
DEV = torch.device("cuda", 0)


def load(name):
    return np.load(os.path.join(GOLD, f"reg_golden_{name}.npz"))


def close(a, b, rtol=LOSS_RTOL):
    assert abs(float(a) - float(b)) <= rtol * max(abs(float(b)), 1e-30), (float(a), float(b))


@pytest.mark.parametrize("name", ["region_a", "region_b"])
def test_region_laplacian_matches_reference_golden(name):
    from sings_b200.regularizers import RegionLaplacianLoss_v2
    z = load(name)
    verts = torch.from_numpy(z["verts"]).to(DEV)
    edges = torch.from_numpy(z["edges"]).to(DEV)
    pos = RegionLaplacianLoss_v2(verts=verts, edges=edges, vertex_labels=z["labels"], region_weights=POSITION_W)   # numpy labels, as gs_trainer.py:180
    col = RegionLaplacianLoss_v2(verts=verts, edges=edges, vertex_labels=torch.from_numpy(z["labels"]).to(DEV),
                                 region_weights=COLOR_W)
    x = torch.from_numpy(z["xyz"]).to(DEV).requires_grad_(True)
    shs = torch.from_numpy(z["shs"]).to(DEV).requires_grad_(True)
    l_pos = pos(x)
    l_col = col(shs[:, 0])                       # (V, 3) view with row stride 48: read in place
    l_hand = pos.forward_hands(x)
    g_pos, = torch.autograd.grad(l_pos, x, retain_graph=True)
    g_hand, = torch.autograd.grad(l_hand, x)
    g_col, = torch.autograd.grad(l_col, shs)
    close(l_pos, z["loss_pos_f64"]); close(l_col, z["loss_col_f64"]); close(l_hand, z["loss_hand_f64"])
    assert_grad_close(g_pos.cpu().numpy(), z["grad_pos_f64"], "region position", tol=GRAD_TOL)
    assert_grad_close(g_hand.cpu().numpy(), z["grad_hand_f64"], "hands", tol=GRAD_TOL)
    assert_grad_close(g_col.cpu().numpy(), z["grad_col_f64"], "region colour", tol=GRAD_TOL)
    assert float(g_col[:, 1:].abs().max()) == 0.0                       # only shs[:, 0] is in the loss
    # the way the trainer uses the terms: weighted sum, one backward (gs_trainer.py:378-397)
    x2 = torch.from_numpy(z["xyz"]).to(DEV).requires_grad_(True)
    total = 1000.0 * 0.5 * pos(x2) + pos.forward_hands(x2) * 1e-5
    total.backward()
    ref = 500.0 * z["grad_pos_f64"] + 1e-5 * z["grad_hand_f64"]
    assert_grad_close(x2.grad.cpu().numpy(), ref, "weighted sum", tol=GRAD_TOL)
    # after densification: rebuilt operator gives the same result (gs_trainer.py:515-521)
    pos.reset_laplacians(verts, edges, torch.from_numpy(z["labels"]).to(DEV))
    close(pos(x.detach()), z["loss_pos_f64"])


def test_pcd_smoothing_matches_reference_golden():
    from sings_b200.regularizers import laplacian, pcd_laplacian_smoothing
    z = load("pcd")
    edges = torch.from_numpy(z["edges"]).to(DEV)
    x = torch.from_numpy(z["pts"]).to(DEV).requires_grad_(True)
    loss = pcd_laplacian_smoothing(x, edges)
    g, = torch.autograd.grad(loss, x)
    close(loss, z["loss_f64"])
    assert_grad_close(g.cpu().numpy(), z["grad_f64"], "pcd", tol=GRAD_TOL)
    # a prebuilt operator and an upstream gradient != 1
    op = laplacian(x, edges)
    x3 = torch.from_numpy(z["pts"]).to(DEV).requires_grad_(True)
    (pcd_laplacian_smoothing(x3, op) * 3.0).backward()
    assert_grad_close(x3.grad.cpu().numpy(), 3.0 * z["grad_f64"], "pcd, prebuilt operator", tol=GRAD_TOL)
    # rows whose Laplacian is exactly 0 get gradient 0, not NaN (torch's norm backward): the zero field
    c = torch.zeros(300, 3, device=DEV, requires_grad=True)
    l0 = pcd_laplacian_smoothing(c, op)
    l0.backward()
    assert float(l0) == 0.0 and float(c.grad.abs().max()) == 0.0 and bool(torch.isfinite(c.grad).all())


@pytest.mark.parametrize("name", ["l2_a", "l2_b", "l2_c"])
def test_l2norm_matches_reference_golden(name):
    from sings_b200.regularizers import L2Norm
    z = load(name)
    c = z["cfg"]
    mod = L2Norm(lambda_xyz_offsets=c[0], lambda_scales_diff=c[1], lambda_max_scale=c[2], max_scale_threshold=c[3],
                 lambda_min_opacity=c[4], min_opacity_threshold=c[5])
    o = torch.from_numpy(z["xyz_offsets"]).to(DEV).requires_grad_(True)
    s = torch.from_numpy(z["scales"]).to(DEV).requires_grad_(True)
    p = torch.from_numpy(z["opacity"]).to(DEV).requires_grad_(True)
    d = {"xyz_offsets": o, "scales": s}
    if bool(z["has_opacity"]):
        d["opacity"] = p
    loss = mod(d)
    (loss * 2.0).backward()
    close(loss, z["loss_f64"])
    assert_grad_close(o.grad.cpu().numpy(), 2.0 * z["grad_off_f64"], "xyz_offsets", tol=GRAD_TOL)
    assert_grad_close(s.grad.cpu().numpy(), 2.0 * z["grad_scales_f64"], "scales", tol=GRAD_TOL)
    assert float(s.grad[:, 1:].abs().max()) == 0.0
    if bool(z["has_opacity"]):
        ref = 2.0 * z["grad_opacity_f64"]
        got = p.grad.cpu().numpy()
        assert got.shape == ref.shape
        if np.abs(ref).max() > 0:
            assert_grad_close(got, ref, "opacity", tol=GRAD_TOL)
        else:
            assert np.abs(got).max() == 0.0                              # nothing below the threshold
    else:
        assert p.grad is None


def test_large_random_graph_against_oracle_properties():
    """110k vertices (the subdivided SMPL template's size), 15 regions, ~4 edges per vertex, C = 1..4, against a
    float64 sparse evaluation with torch on the host; plus linearity in the upstream gradient."""
    from sings_b200.regularizers import laplacian_loss, region_laplacian
    g = torch.Generator().manual_seed(3)
    V, R_ = 110_000, 15
    labels = torch.sort(torch.randint(0, R_, (V,), generator=g)).values
    a = torch.arange(V).repeat_interleave(2)
    b = (a + torch.randint(1, 40, (2 * V,), generator=g)).clamp(max=V - 1)
    edges = torch.unique(torch.sort(torch.stack([a, b], 1), dim=1).values, dim=0)
    edges = edges[edges[:, 0] != edges[:, 1]]
    op, n_region = region_laplacian(labels.to(DEV), edges.to(DEV))
    w_lab = torch.rand(R_, generator=g, dtype=torch.float64)
    rp, ci, va = op.row_ptr.cpu().long(), op.col_idx.cpu().long(), op.vals.cpu().double()
    rows = torch.repeat_interleave(torch.arange(V), rp[1:] - rp[:-1])
    for C in (1, 2, 3, 4):
        row_w64 = w_lab[labels] / (n_region.cpu().double()[labels] * C)
        row_w = row_w64.to(torch.float32).to(DEV)
        xh = torch.randn(V, C, generator=g)
        for mode in (0, 1):
            x64 = xh.double().requires_grad_(True)
            y = torch.zeros(V, C, dtype=torch.float64).index_add(0, rows, va[:, None] * x64[ci])
            ref = (row_w64 * ((y ** 2).sum(1) if mode == 0 else y.norm(dim=1))).sum()
            gref, = torch.autograd.grad(ref, x64)
            x = xh.to(DEV).requires_grad_(True)
            loss = laplacian_loss(op, x, row_w, mode)
            (loss * 0.25).backward()
            close(loss, ref, 2e-5)
            assert_grad_close(x.grad.cpu().numpy(), 0.25 * gref.numpy(), f"C={C} mode={mode}", tol=GRAD_TOL)


def test_build_edges_and_smoothing_module():
    """build_edges on the library's K-NN == the brute-force restatement (as sets per point: equidistant ties may
    be ordered differently), and LaplacianSmoothing over a dict == the oracle's sum."""
    from oracle import reg_oracle as ro
    from sings_b200.regularizers import LaplacianSmoothing, build_edges
    g = torch.Generator().manual_seed(9)
    N, K = 2000, 9
    pts = torch.rand(N, 3, generator=g)
    e = build_edges(pts.to(DEV), K).cpu()
    e_ref = ro.build_edges(pts, K)
    assert e.shape == e_ref.shape == (N * K, 2) and e.dtype == torch.int64
    assert torch.equal(e[:, 0], e_ref[:, 0])
    assert torch.equal(torch.sort(e[:, 1].reshape(N, K), dim=1).values, torch.sort(e_ref[:, 1].reshape(N, K), dim=1).values)
    d = {"xyz_canon": pts.to(DEV).requires_grad_(True), "color": torch.rand(N, 3, generator=g).to(DEV).requires_grad_(True)}
    loss = LaplacianSmoothing(K=K)(d)
    loss.backward()
    ref = 0.0
    grads = {}
    for k, v in d.items():
        x64 = v.detach().cpu().double().requires_grad_(True)
        l = ro.pcd_laplacian_smoothing(x64, e_ref)
        grads[k], = torch.autograd.grad(l, x64)
        ref = ref + float(l)
    close(loss, ref, 2e-5)
    for k, v in d.items():
        assert_grad_close(v.grad.cpu().numpy(), grads[k].numpy(), k, tol=GRAD_TOL)


def test_l2norm_full_size_and_edge_cases():
    from oracle import reg_oracle as ro
    from sings_b200.regularizers import L2Norm
    g = torch.Generator().manual_seed(4)
    N = 200_000
    off = 0.01 * torch.randn(N, 3, generator=g)
    sc = torch.exp(torch.log(torch.tensor(0.002)) + torch.rand(N, 1, generator=g) * 1.8).repeat(1, 3)
    op = torch.sigmoid(1.5 * torch.randn(N, 1, generator=g))
    cfg = dict(lambda_xyz_offsets=0.001, lambda_scales_diff=0.005, max_scale_threshold=0.005, lambda_max_scale=0.01,
               min_opacity_threshold=0.2, lambda_min_opacity=0.001)
    t64 = [t.double().requires_grad_(True) for t in (off, sc, op)]
    ref = ro.l2norm({"xyz_offsets": t64[0], "scales": t64[1], "opacity": t64[2]}, **cfg)
    gref = torch.autograd.grad(ref, t64)
    t = [x.to(DEV).requires_grad_(True) for x in (off, sc, op)]
    loss = L2Norm(**cfg)({"xyz_offsets": t[0], "scales": t[1], "opacity": t[2]})
    loss.backward()
    close(loss, ref)
    for a, b, nm in zip(t, gref, ("xyz_offsets", "scales", "opacity")):
        assert_grad_close(a.grad.cpu().numpy(), b.numpy(), nm, tol=GRAD_TOL)
    # all-zero offsets, constant scales: the norms are 0 and so are their gradients (no NaN)
    z = {"xyz_offsets": torch.zeros(100, 3, device=DEV, requires_grad=True),
         "scales": torch.full((100, 3), 0.004, device=DEV, requires_grad=True)}
    l = L2Norm(**cfg)(z)
    l.backward()
    assert abs(float(l)) < 1e-9          # (sum s^2 - (sum s)^2 / N in binary64 may leave ~1e-22 under the root)
    assert float(z["xyz_offsets"].grad.abs().max()) == 0.0 and float(z["scales"].grad.abs().max()) == 0.0
    # only the scales need a gradient
    s_only = sc.to(DEV).requires_grad_(True)
    L2Norm(**cfg)({"xyz_offsets": off.to(DEV), "scales": s_only}).backward()
    assert s_only.grad is not None and bool(torch.isfinite(s_only.grad).all())

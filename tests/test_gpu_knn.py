"""Exact K-NN + scale-edge loss (sings_b200/csrc/knn.cu through the C ABI) against the brute-force
oracle (oracle/knn_oracle.py; pytorch3d itself is not installed: parity unpinned for knn_points,
see the oracle's header).  Neighbour SETS must be exact; squared distances within 1e-6 relative
(binary32 sums in another order); indices equal wherever distances are not tied."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def brute(x, K):
    from oracle import knn_oracle as ko
    d, idx, _ = ko.knn_points(x[None], x[None], K + 1)
    return d[0, :, 1:], idx[0, :, 1:]


def clouds():
    g = torch.Generator().manual_seed(4)
    from sings_b200 import synthetic as syn
    av = syn.make_avatar(6000, 24, seed=2)
    yield "avatar", torch.from_numpy(av.xyz_canon.astype(np.float32))                       # surface-like
    yield "volume", torch.rand(5000, 3, generator=g) * torch.tensor([2.0, 0.5, 1.0]) - 0.7   # volumetric, anisotropic box
    yield "clustered", torch.cat([torch.randn(3000, 3, generator=g) * 0.01, torch.randn(500, 3, generator=g) * 3.0])
    yield "planar", torch.cat([torch.rand(4000, 2, generator=g), torch.zeros(4000, 1)], 1)  # zero extent along z
    yield "tiny", torch.rand(5, 3, generator=g)                                              # fewer than K + 1 points


@pytest.mark.parametrize("name,x", list(clouds()), ids=[n for n, _ in clouds()])
def test_knn_exact(name, x):
    from sings_b200.losses import knn_points
    K = 8
    dev = torch.device("cuda", 0)
    mean, idx, d2 = knn_points(x.to(dev), K, return_index=True)
    mean, idx, d2 = mean.cpu(), idx.cpu().long(), d2.cpu()
    N = x.shape[0]
    Ke = min(K, N - 1)
    dref, iref = brute(x, Ke)
    assert torch.allclose(d2[:, :Ke].double(), dref.double(), rtol=2e-6, atol=1e-12)
    if Ke < K:
        assert bool((idx[:, Ke:] == -1).all())
    # indices: equal except inside groups of (nearly) tied distances
    diff = idx[:, :Ke] != iref
    if diff.any():
        r, c = diff.nonzero(as_tuple=True)
        mine = ((x[idx[r, c]] - x[r]) ** 2).sum(1)
        assert torch.allclose(mine.double(), dref[r, c].double(), rtol=2e-6, atol=1e-12)
    ref_mean = dref.double().sqrt().mean(1)
    assert torch.allclose(mean.double(), ref_mean, rtol=1e-5, atol=1e-9)


def test_edge_loss_matches_the_reference_formula_and_is_differentiable():
    from oracle import knn_oracle as ko
    from sings_b200 import synthetic as syn
    from sings_b200.losses import GaussiansEdgeLoss
    av = syn.make_avatar(8000, 24, seed=5, isotropic=True)
    xyz, sc = torch.from_numpy(av.xyz_canon.astype(np.float32)), torch.from_numpy(av.scales.astype(np.float32))
    so = sc.clone().requires_grad_(True)
    lo = ko.gaussians_edge_loss({"xyz_canon": xyz, "scales": so}, K=9)
    lo.backward()
    dev = torch.device("cuda", 0)
    sg = sc.to(dev).requires_grad_(True)
    lg = GaussiansEdgeLoss(K=9)({"xyz_canon": xyz.to(dev), "scales": sg})
    lg.backward()
    assert abs(float(lg.detach()) - float(lo.detach())) <= 1e-5 * abs(float(lo.detach()))
    assert torch.allclose(sg.grad.cpu(), so.grad, rtol=1e-4, atol=1e-10)


def test_full_size_properties():
    """200k points (BASELINE config c2): every reported neighbour distance is reproduced from the
    indices, rows ascend, no point is its own neighbour, and a random sample of rows is exact."""
    from sings_b200 import synthetic as syn
    from sings_b200.losses import knn_points
    av = syn.make_avatar(200_000, 24, seed=0)
    x = torch.from_numpy(av.xyz_canon.astype(np.float32))
    dev = torch.device("cuda", 0)
    xd = x.to(dev)
    mean, idx, d2 = knn_points(xd, 8, return_index=True)
    torch.cuda.synchronize()
    assert bool((idx >= 0).all()) and bool((idx != torch.arange(x.shape[0], device=dev)[:, None]).all())
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    rec = ((xd[idx.long()] - xd[:, None]) ** 2).sum(-1)
    assert torch.allclose(rec, d2, rtol=2e-6, atol=1e-12)
    rows = torch.randperm(x.shape[0], generator=torch.Generator().manual_seed(1))[:400]
    d = torch.cdist(x[rows].double(), x.double()) ** 2
    d[torch.arange(400), rows] = float("inf")
    dref = torch.topk(d, 8, dim=1, largest=False, sorted=True)[0]
    assert torch.allclose(d2[rows.to(dev)].cpu().double(), dref, rtol=2e-6, atol=1e-12)
    assert torch.allclose(mean[rows.to(dev)].cpu().double(), dref.sqrt().mean(1), rtol=1e-5)

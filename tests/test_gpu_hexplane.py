"""Fused tri-plane interpolation (sings_b200/csrc/hexplane.cu through the C ABI and the HexPlaneField
mirror) against the reference-generated golden vectors and the CPU oracle.  Floating point: features
and gradients within 1e-5 of the tensor's largest element (binary32 sums in another order)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def close(a, b, name):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = float((a - b).abs().max())
    assert err <= TOL * max(float(b.abs().max()), 1e-12), f"{name}: max abs error {err} vs scale {float(b.abs().max())}"


def field_from(z, dev):
    from sings_b200.triplane import HexPlaneField
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": int(z["C"]),
           "resolution": [int(v) for v in z["reso"]], "multires": [int(v) for v in z["multires"]]}
    f = HexPlaneField(cfg, bounds=float(z["bounds"]), device=dev)
    params = [p for gp in f.grids for p in gp]
    with torch.no_grad():
        for i, p in enumerate(params):
            assert tuple(p.shape) == tuple(z[f"plane_{i}"].shape)
            p.copy_(torch.from_numpy(z[f"plane_{i}"]))
    return f, params


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "hexplane_golden_*.npz"))))
def test_matches_reference_golden(path):
    z = np.load(path)
    dev = torch.device("cuda", 0)
    f, params = field_from(z, dev)
    assert np.allclose(f.aabb.cpu().numpy(), z["aabb"])
    pts = torch.from_numpy(z["pts"]).to(dev).requires_grad_(True)
    feats = f(pts)
    close(feats, torch.from_numpy(z["feats"]), "features")
    grads = torch.autograd.grad((feats * torch.from_numpy(z["d_out"]).to(dev)).sum(), [pts] + params)
    close(grads[0], torch.from_numpy(z["d_pts"]), "d_pts")
    for i, g in enumerate(grads[1:]):
        close(g, torch.from_numpy(z[f"d_plane_{i}"]), f"d_plane_{i}")


def test_shipped_configuration_against_oracle():
    """human_complex.yaml:39-43: 32 channels, 64^3, multires [1, 2, 4]; 20k points; state_dict keys of the reference."""
    from oracle import hexplane_oracle as ho
    from sings_b200.triplane import HexPlaneField
    dev = torch.device("cuda", 0)
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": 32, "resolution": [64, 64, 64],
           "multires": [1, 2, 4]}
    torch.manual_seed(3)
    f = HexPlaneField(cfg, bounds=1.3, device=dev)
    assert f.feat_dim == 96
    assert sorted(f.state_dict().keys()) == sorted([f"grids.{s}.{p}" for s in range(3) for p in range(3)])
    g = torch.Generator().manual_seed(7)
    pts = (torch.rand(20000, 3, generator=g) * 2 - 1) * 1.4
    d_out = torch.randn(20000, 96, generator=g)
    params = [p for gp in f.grids for p in gp]
    pg = pts.to(dev).requires_grad_(True)
    feats = f(pg)
    grads = torch.autograd.grad((feats * d_out.to(dev)).sum(), [pg] + params)
    cpu_planes = [p.detach().cpu().clone().requires_grad_(True) for p in params]
    pc = pts.clone().requires_grad_(True)
    fo = ho.hexplane_features(pc, f.aabb.detach().cpu(), [cpu_planes[3 * s:3 * s + 3] for s in range(3)])
    go = torch.autograd.grad((fo * d_out).sum(), [pc] + cpu_planes)
    close(feats, fo, "features")
    close(grads[0], go[0], "d_pts")
    for i in range(9):
        close(grads[1 + i], go[1 + i], f"d_plane_{i}")
    with pytest.raises(Exception):
        f(pts)                      # CPU tensor: no CPU path

"""The tri-plane oracle against golden vectors made by the reference's own HexPlaneField
(tests/golden/make_hexplane_golden.py; hexplane.py:18-189)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hexplane_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(path):
    z = np.load(path)
    S = len(z["multires"])
    planes = [torch.from_numpy(z[f"plane_{i}"]) for i in range(3 * S)]
    grids = [planes[3 * s:3 * s + 3] for s in range(S)]
    return z, grids, planes


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "hexplane_golden_*.npz"))))
def test_oracle_matches_reference_golden(path):
    z, grids, planes = load(path)
    for p in planes:
        p.requires_grad_(True)
    pts = torch.from_numpy(z["pts"]).requires_grad_(True)
    feats = ho.hexplane_features(pts, torch.from_numpy(z["aabb"]), grids)
    assert torch.equal(feats.detach(), torch.from_numpy(z["feats"]))
    grads = torch.autograd.grad((feats * torch.from_numpy(z["d_out"])).sum(), [pts] + planes)
    assert torch.allclose(grads[0], torch.from_numpy(z["d_pts"]), rtol=0, atol=0)
    for i, g in enumerate(grads[1:]):
        assert torch.allclose(g, torch.from_numpy(z[f"d_plane_{i}"]), rtol=1e-6, atol=1e-7)     # (scatter-add order)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "hexplane_golden_*.npz"))) == 2

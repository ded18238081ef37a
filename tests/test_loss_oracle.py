"""The image-loss oracle against golden vectors made by the reference's own l1_loss / ssim
(tests/golden/make_loss_golden.py; utils.py:16-70, loss.py:57-70)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(path):
    z = np.load(path)
    gt = torch.from_numpy(z["gt_u8"].astype(np.float32) / 255.0).permute(2, 0, 1).contiguous()
    mask = torch.from_numpy(z["mask"]) if z["mask"].size else None
    return z, torch.from_numpy(z["pred"]), gt, mask, torch.from_numpy(z["bg"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "loss_golden_*_f*.npz"))))
def test_oracle_matches_reference_golden(path):
    z, pred, gt, mask, bg = load(path)
    dt = torch.float64 if path.endswith("f64.npz") else torch.float32
    tol = 1e-12 if dt == torch.float64 else 2e-6
    p = pred.to(dt).requires_grad_(True)
    loss, items, _ = lo.human_image_loss(p, gt.to(dt), None if mask is None else mask.to(dt), bg.to(dt))
    (g,) = torch.autograd.grad(loss, p)
    assert abs(float(loss.detach()) - float(z["loss"])) <= tol * max(1.0, abs(float(z["loss"])))
    assert abs(float(items["l1"]) - float(z["l1"])) <= tol * max(1.0, abs(float(z["l1"])))
    assert abs(float(items["ssim"]) - float(z["ssim"])) <= tol * max(1.0, abs(float(z["ssim"])))
    ref = torch.from_numpy(z["grad"]).to(dt)
    assert float((g - ref).abs().max()) <= tol * float(ref.abs().max()) * 10


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "loss_golden_*_f32.npz"))) == 3

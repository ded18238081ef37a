"""Generate golden vectors for the image loss by running the REFERENCE's own functions on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_loss_golden.py
Writes tests/golden/loss_golden_<case>.npz.  Executed reference code:
  l1_loss, ssim (create_window, gaussian, _ssim)   /root/reference/sings/rec/losses/utils.py:16-70
composed by the statements of HumanLoss.forward    /root/reference/sings/rec/losses/loss.py:57-70, 88-90
(loss.py itself cannot be imported here: it needs lpips and a CUDA device at construction;
utils.py imports pytorch3d and a body-model parser at module level, neither of which the image
terms use -- empty stand-in modules satisfy those imports).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_utils():
    for name in ("pytorch3d", "pytorch3d.ops", "sings", "sings.rec", "sings.rec.utils", "sings.rec.utils.body_model",
                 "sings.rec.utils.body_model.smpl_parsing"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ops = sys.modules["pytorch3d.ops"]
    for n in ("knn_points", "laplacian", "cot_laplacian", "norm_laplacian"):
        setattr(ops, n, None)
    sys.modules["sings.rec.utils.body_model.smpl_parsing"].parse_weights = None
    spec = importlib.util.spec_from_file_location("ref_loss_utils", f"{REF}/sings/rec/losses/utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    U = load_reference_utils()
    cases = [dict(name="a", H=45, W=70, masked=True, seed=1), dict(name="b", H=64, W=64, masked=False, seed=2),
             dict(name="c", H=33, W=17, masked=True, seed=3)]
    for c in cases:
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            g = torch.Generator().manual_seed(c["seed"])
            H, W = c["H"], c["W"]
            # smooth-ish images in [0, 1] with flat regions (SSIM denominators near C1, C2 matter there)
            base = torch.rand(3, H // 4 + 2, W // 4 + 2, generator=g)
            gt = torch.nn.functional.interpolate(base[None], size=(H, W), mode="bilinear", align_corners=False)[0]
            pred = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1.2)
            gt_u8 = (gt.clamp(0, 1) * 255).round().to(torch.uint8)          # what a dataset holds
            gt = gt_u8.to(torch.float32) / 255.0
            mask = None
            if c["masked"]:
                yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
                mask = (((yy - H / 2) / (H / 2.5)) ** 2 + ((xx - W / 2) / (W / 3.0)) ** 2 < 1.0).to(torch.float32)
            bg = torch.tensor([1.0, 1.0, 1.0]) if c["name"] != "c" else torch.tensor([0.1, 0.5, 0.9])
            pred_l = pred.to(dt).clone().requires_grad_(True)
            gt_d, bg_d = gt.to(dt), bg.to(dt)
            m = torch.ones(1, H, W, dtype=dt) if mask is None else mask.to(dt).unsqueeze(0)      # data['mask'].unsqueeze(0)
            # ---- HumanLoss.forward, loss.py:57-70 and :88-90, l_l1_w = 0.8, l_ssim_w = 0.2
            gt_image = gt_d * m + bg_d[:, None, None] * (1.0 - m)
            Ll1 = U.l1_loss(pred_l, gt_image, m)
            l1_item = 0.8 * Ll1
            loss_ssim = 1.0 - U.ssim(pred_l, gt_image)
            loss_ssim = loss_ssim * (m.sum() / (pred_l.shape[-1] * pred_l.shape[-2]))
            ssim_item = 0.2 * loss_ssim
            loss = 0.0
            for v in (l1_item, ssim_item):
                loss += v
            (grad,) = torch.autograd.grad(loss, pred_l)
            np.savez_compressed(
                os.path.join(HERE, f"loss_golden_{c['name']}_{tag}.npz"),
                pred=pred.numpy(), gt_u8=gt_u8.permute(1, 2, 0).contiguous().numpy(),
                mask=(np.zeros(0, np.float32) if mask is None else mask.numpy()), bg=bg.numpy(),
                loss=loss.detach().numpy(), l1=l1_item.detach().numpy(), ssim=ssim_item.detach().numpy(),
                grad=grad.numpy())
            print(c["name"], tag, float(loss), float(l1_item), float(ssim_item))


if __name__ == "__main__":
    main()

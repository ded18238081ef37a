"""CPU restatement of the reference's image loss (TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's comparison legs may import it; the product path never does).

Follows, statement by statement,
  /root/reference/sings/rec/losses/utils.py:16-20   l1_loss
  /root/reference/sings/rec/losses/utils.py:27-37   gaussian, create_window (11 taps, sigma 1.5)
  /root/reference/sings/rec/losses/utils.py:40-70   ssim, _ssim (zero padding, C1 = 0.01^2, C2 = 0.03^2)
  /root/reference/sings/rec/losses/loss.py:57-70    HumanLoss.forward: target compositing, the two terms
PINNED: tests/golden/loss_golden_*.npz hold outputs and gradients of the reference's own functions
(tests/golden/make_loss_golden.py); tests/test_loss_oracle.py checks this file against them.
"""
from math import exp

import torch
import torch.nn.functional as F


def gaussian(window_size: int, sigma: float) -> torch.Tensor:
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def create_window(window_size: int, channel: int, dtype) -> torch.Tensor:
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous().to(dtype)


def l1_loss(pred, gt, mask=None):
    if mask is not None:
        return torch.abs(pred - gt).sum() / mask.sum()
    return torch.abs(pred - gt).mean()


def ssim(img1, img2, window_size: int = 11):
    channel = img1.size(-3)
    window = create_window(window_size, channel, img1.dtype).to(img1.device)     # utils.py:44-46
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def human_image_loss(pred, gt, mask, bg_color, l_l1_w: float = 0.8, l_ssim_w: float = 0.2):
    """loss.py:57-70 (+ the sum of :88-90) for the L1 and SSIM terms.  pred, gt (3,H,W); mask (H,W)
    or None; bg_color (3,).  Returns (loss, {'l1', 'ssim'}, composited target)."""
    H, W = pred.shape[-2:]
    m = torch.ones(1, H, W, dtype=pred.dtype, device=pred.device) if mask is None else mask.reshape(1, H, W).to(pred.dtype)
    gt_c = gt * m + bg_color[:, None, None] * (1.0 - m)
    items = {"l1": l_l1_w * l1_loss(pred, gt_c, m)}
    loss_ssim = (1.0 - ssim(pred, gt_c)) * (m.sum() / (W * H))
    items["ssim"] = l_ssim_w * loss_ssim
    return items["l1"] + items["ssim"], items, gt_c

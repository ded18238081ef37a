"""Image loss of a training step on the device: fused L1 + SSIM.

Host-side mirror of the reference's HumanLoss for its two image terms
(/root/reference/sings/rec/losses/loss.py:15-92 with l_lpips_w = 0; l1_loss / ssim of
/root/reference/sings/rec/losses/utils.py:16-70): same argument meaning, same weights
(l_l1_w = 0.8, l_ssim_w = 0.2), same loss_dict keys.  One forward and one backward kernel
(sings_b200/csrc/image_loss.cu) instead of five conv2d launches, ~25 elementwise kernels and their
autograd graph; the ground truth may stay the dataset's uint8 (H, W, 3) image.
There is no CPU path: CPU tensors raise SgsError.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import SgsError, raw_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, mask, bg, w_l1: float, w_ssim: float):
        if not pred.is_cuda:
            raise SgsError("sings_b200.losses needs CUDA tensors (there is no CPU path)")
        if pred.dim() != 3 or pred.shape[0] != 3 or pred.dtype != torch.float32:
            raise SgsError("pred must be a (3, H, W) float32 image")
        _, H, W = pred.shape
        pred = pred.contiguous()
        u8 = gt.dtype == torch.uint8
        if u8:
            if tuple(gt.shape) != (H, W, 3):
                raise SgsError("a uint8 ground truth must be (H, W, 3)")
        elif tuple(gt.shape) != (3, H, W) or gt.dtype != torch.float32:
            raise SgsError("gt must be (3, H, W) float32 or (H, W, 3) uint8")
        gt = gt.contiguous()
        if mask is not None:
            mask = mask.reshape(H, W).to(torch.float32).contiguous()
        bg = bg.to(device=pred.device, dtype=torch.float32).contiguous()
        L = _lib.lib()
        scratch = torch.empty(L.sgs_image_loss_scratch_floats(H, W), device=pred.device, dtype=torch.float32)
        sums = torch.empty(4, device=pred.device, dtype=torch.float64)
        loss3 = torch.empty(3, device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            _lib.check(L.sgs_image_loss_fwd(H, W, _p(pred), _p(gt), int(u8), _p(mask), _p(bg), _p(scratch), _p(sums),
                                            float(w_l1), float(w_ssim), _p(loss3), raw_stream(pred.device)),
                       "sgs_image_loss_fwd")
        ctx.save_for_backward(pred, scratch, sums)
        ctx.hw, ctx.w = (H, W), (float(w_l1), float(w_ssim))
        loss, l1, ss = loss3[0].clone(), loss3[1].clone(), loss3[2].clone()      # (not views of one buffer: autograd outputs)
        ctx.mark_non_differentiable(l1, ss)
        return loss, l1, ss

    @staticmethod
    def backward(ctx, dloss, _dl1, _dssim):
        pred, scratch, sums = ctx.saved_tensors
        H, W = ctx.hw
        out = torch.empty_like(pred)
        dl = dloss.to(torch.float32).contiguous()
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().sgs_image_loss_bwd(H, W, _p(pred), _p(scratch), _p(sums), ctx.w[0], ctx.w[1], _p(dl),
                                                     _p(out), None, raw_stream(pred.device)), "sgs_image_loss_bwd")
        return out, None, None, None, None, None


def image_loss(pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor], bg_color: torch.Tensor,
               l_l1_w: float = 0.8, l_ssim_w: float = 0.2) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """loss, loss_dict of HumanLoss.forward (loss.py:41-92) for the L1 and SSIM terms:
    loss_dict['l1'] = l_l1_w * sum|pred - gt'| / sum(mask), loss_dict['ssim'] = l_ssim_w *
    (1 - ssim(pred, gt')) * (sum(mask) / (H W)), gt' = gt * mask + bg_color * (1 - mask).
    pred (3, H, W) float32 (differentiable); gt (3, H, W) float32 or (H, W, 3) uint8; mask (H, W) /
    (1, H, W) or None (= ones); bg_color (3,)."""
    loss, l1, ss = _ImageLoss.apply(pred, gt, mask, bg_color, l_l1_w, l_ssim_w)
    return loss, {"l1": l_l1_w * l1, "ssim": l_ssim_w * ss}


class ImageLossBuffers:
    """Allocation-free form for a captured frame (what bench.py's end-to-end loop uses): the ground
    truth is uploaded as uint8, the loss gradient lands in a preallocated dL/dimage."""

    def __init__(self, H: int, W: int, device, l_l1_w: float = 0.8, l_ssim_w: float = 0.2):
        self.H, self.W, self.dev = H, W, torch.device(device)
        self.w = (float(l_l1_w), float(l_ssim_w))
        L = _lib.lib()
        self.scratch = torch.empty(L.sgs_image_loss_scratch_floats(H, W), device=self.dev, dtype=torch.float32)
        self.sums = torch.empty(4, device=self.dev, dtype=torch.float64)
        self.dL_dimage = torch.empty(3, H, W, device=self.dev, dtype=torch.float32)
        self.loss_value = torch.zeros(1, device=self.dev, dtype=torch.float32)

    def run(self, pred: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor], bg: torch.Tensor, stream=None):
        """forward + backward of the loss (dloss = 1): returns dL/dimage; the loss lands in
        self.loss_value, the three sums stay in self.sums."""
        L = _lib.lib()
        st = raw_stream(self.dev) if stream is None else stream
        _lib.check(L.sgs_image_loss_fwd(self.H, self.W, _p(pred), _p(gt), int(gt.dtype == torch.uint8), _p(mask), _p(bg),
                                        _p(self.scratch), _p(self.sums), self.w[0], self.w[1], None, st), "sgs_image_loss_fwd")
        _lib.check(L.sgs_image_loss_bwd(self.H, self.W, _p(pred), _p(self.scratch), _p(self.sums), self.w[0], self.w[1],
                                        None, _p(self.dL_dimage), _p(self.loss_value), st), "sgs_image_loss_bwd")
        return self.dL_dimage

    def loss(self) -> torch.Tensor:
        hw = float(self.H * self.W)
        s = self.sums
        return (self.w[0] * s[0] / s[2] + self.w[1] * (1.0 - s[1] / (3.0 * hw)) * (s[2] / hw)).to(torch.float32)


# ------------------------------------------------------------------------------------------
# neighbour distances and the scale-edge loss (loss_items.py:57-90)
# ------------------------------------------------------------------------------------------
def knn_points(xyz: torch.Tensor, K: int = 8, return_index: bool = False):
    """Exact K nearest neighbours of every point of xyz (N, 3) among the OTHER points of the set
    (sgs_knn_mean_dist): returns mean_dist (N,) = mean |x_j - x_i| over the K neighbours, and with
    return_index also idx (N, K) int32 nearest first and dist2 (N, K) ascending -- the columns 1..K
    of pytorch3d.ops.knn_points(xyz[None], xyz[None], K=K + 1), whose column 0 is the point itself.
    No gradient flows through it (the reference detaches the lengths, loss_items.py:79)."""
    if not xyz.is_cuda:
        raise SgsError("sings_b200.losses needs CUDA tensors (there is no CPU path)")
    x = xyz.detach().to(torch.float32).contiguous()
    N = x.shape[0]
    L = _lib.lib()
    scratch = torch.empty(int(L.sgs_knn_scratch_bytes(N)) + 256, device=x.device, dtype=torch.uint8)
    base = scratch.data_ptr()
    off = (-base) % 256
    mean = torch.empty(N, device=x.device, dtype=torch.float32)
    idx = torch.empty(N, K, device=x.device, dtype=torch.int32) if return_index else None
    d2 = torch.empty(N, K, device=x.device, dtype=torch.float32) if return_index else None
    with torch.cuda.device(x.device):
        _lib.check(L.sgs_knn_mean_dist(N, _p(x), int(K), base + off, scratch.numel() - off, _p(mean), _p(idx), _p(d2),
                                       raw_stream(x.device)), "sgs_knn_mean_dist")
    return (mean, idx, d2) if return_index else mean


class GaussiansEdgeLoss(torch.nn.Module):
    """loss_items.py:57-90, same constructor and call: `loss = GaussiansEdgeLoss(K=9)(human_gs_out)` with
    human_gs_out['xyz_canon'] (N, 3) and human_gs_out['scales'] (N, 3) (isotropic: column 0 is used).
    K counts the point itself, as in the reference (K=9 -> 8 neighbours).  The neighbour search -- the
    expensive part, run over every Gaussian every iteration -- is one call of the library; the loss
    ((scale_i - mean edge length_i)^2).mean() and its gradient to the scales stay torch ops."""

    def __init__(self, K: int = 9, eps: float = 1e-12):
        super().__init__()
        self._K, self._eps = K, eps

    def forward(self, human_gs_out):
        verts = human_gs_out["xyz_canon"]
        scales = human_gs_out["scales"][:, 0]
        edge_lengths = knn_points(verts, self._K - 1).unsqueeze(1)          # (N, 1), detached
        scale_proj_i = scales.unsqueeze(1)
        len_factor = 1.0
        return ((scale_proj_i - len_factor * edge_lengths) ** 2).mean()

"""Forward-only rendering of a pose sequence (the reference's animation path), frames sharded
over ranks.

Mirrors /root/reference/sings/rec/trainer/gs_trainer.py:664-728 (`animate_chunk`: per frame
`human_gs.forward` under `torch.no_grad()` then `render_human_scene`) for the part that is on
the hot path: pose -> A -> LBS -> rasterize.  Frames are independent given the replicated
canonical Gaussians, so rank r of R renders the contiguous range `dp.shard_frames(F, r, R)`
with no collective (BASELINE.json configs[3]).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import dp
from .step import AvatarStep, FrameInputs


def frame_to_uint8(img: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(3,H,W) float image -> (H,W,3) uint8 on the device: the reference's clamp(0,1)
    (gs_renderer_single.py:96) followed by the 8-bit conversion its writers apply on the host after
    `.cpu()` (gs_trainer.py:716-719: `(image * 255).astype(uint8)` for cv2.imwrite) -- done before
    the device -> host copy, which then moves a quarter of the bytes."""
    q = torch.mul(img.clamp(0.0, 1.0), 255.0).to(torch.uint8).permute(1, 2, 0)
    if out is None:
        return q.contiguous()
    out.copy_(q)
    return out


def render_frames(step: AvatarStep, frames: Sequence[FrameInputs], rank: int = 0, world: int = 1,
                  out: Optional[torch.Tensor] = None, clamp: bool = True) -> Tuple[int, int, torch.Tensor]:
    """Render this rank's share of `frames`; returns (lo, hi, images (hi-lo, 3, H, W)).

    `out` (optional, (>= hi-lo, 3, H, W) on the device) receives the images; `clamp` applies
    the reference's `torch.clamp(rendered_image, 0, 1)` (gs_renderer_single.py:96).  The
    launch sequence per frame has no host synchronisation; the pair-list capacity is checked
    once at the end (and the affected frames re-rendered if it had to grow)."""
    lo, hi = dp.shard_frames(len(frames), rank, world)
    n = hi - lo
    if out is None:
        out = torch.empty(max(n, 0), 3, step.H, step.Wd, device=step.dev, dtype=torch.float32)
    from ._lib import SgsError
    first = lo
    while first < hi:
        overflow_at = None
        for f in range(first, hi):
            # every frame of a block reports {num_rendered, overflow} into its own pinned row, so one
            # stream sync per block of COUNTER_SLOTS frames sees the flag of each of them
            img = step.forward(frames[f], slot=(f - first) % step.COUNTER_SLOTS)
            out[f - lo].copy_(img)
            if (f - first) % step.COUNTER_SLOTS == step.COUNTER_SLOTS - 1 or f == hi - 1:
                torch.cuda.current_stream(step.dev).synchronize()
                try:
                    step.check_capacity()
                except SgsError:
                    overflow_at = first + (f - first) // step.COUNTER_SLOTS * step.COUNTER_SLOTS
                    break
        if overflow_at is None:
            break
        first = overflow_at            # capacity has been raised: redo the last block
    if clamp:
        out[:n].clamp_(0.0, 1.0)
    return lo, hi, out[:n]

"""Data-parallel sharding of views / frames and the one exchange step of the hot path.

The reference is single-process (no distributed code at all; one view per
step, /root/reference/sings/rec/trainer/gs_trainer.py:214-244; frames independent in
animate_chunk, :681-719).  A frame is an independent unit of work given replicated canonical
Gaussians, so the path shards naturally:

  * training step: view v -> rank v mod R; every rank runs the full hot path on its views;
    ONE all-reduce(SUM) of the flat gradient bucket (canonical-parameter gradients +
    this step's `xyz_gradient_accum` / `denom` increments, whose norms were taken per view
    before the sum -- sings_hybrid.py:1013-1015) and ONE all-reduce(MAX) of `max_radii2D`
    (gs_trainer.py:487-490);
  * animation: contiguous frame ranges per rank, no collective per frame.

One process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Training views of this rank: v with v mod world == rank."""
    return list(range(rank, n_views, world))


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous animation frame range [lo, hi) of this rank (sizes differ by at most 1)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradExchange:
    """Owns the persistent densification statistics and performs the per-step exchange.

    bucket: flat float32 tensor [param grads (n_param_grads) | accum increment (N) | denom
    increment (N)]; max_radii: (N) this step's per-rank maximum screen radius.
    """

    def __init__(self, n_gaussians: int, n_param_grads: int, device, group=None,
                 average_grads: bool = False, defer_max: bool = False):
        # defer_max: max is idempotent and commutes with the running maximum over steps, so the
        # cross-rank MAX of max_radii2D can be taken once when the statistics are consumed
        # (sync_max(), at densification time) instead of every step: one collective per step.
        self.defer_max = defer_max
        self.N = n_gaussians
        self.n_param_grads = n_param_grads
        self.group = group
        self.average = average_grads
        z = lambda: torch.zeros(n_gaussians, device=device, dtype=torch.float32)
        self.xyz_gradient_accum, self.denom, self.max_radii2D = z(), z(), z()

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def exchange(self, bucket: torch.Tensor, max_radii: torch.Tensor, async_op: bool = False,
                 reset_step: bool = True):
        """All-reduce one step's bucket (SUM) and radii (MAX) in place and fold the statistics
        into the persistent accumulators.  With async_op=True returns a finish() callable so
        the exchange overlaps whatever the caller launches next.  reset_step (default) also
        clears the step's statistics (the bucket's accum / denom slices and max_radii) for the
        next view; on CUDA the fold and the clearing are one kernel (sgs_fold_stats).
        AvatarStep ACCUMULATES its statistics into those slices (+=), so a caller that passes
        reset_step=False must call AvatarStep.reset_stats() itself before the next view --
        otherwise the next exchange folds the same increments into the accumulators again."""
        handles = []
        if self.world > 1:
            handles.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if not self.defer_max:
                handles.append(dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=self.group, async_op=True))

        def finish():
            for h in handles:
                h.wait()
            n, N = self.n_param_grads, self.N
            if self.average and self.world > 1:
                bucket[:n].div_(self.world)
            s_acc, s_den = bucket[n:n + N], bucket[n + N:n + 2 * N]
            if bucket.is_cuda and reset_step:
                from . import _lib
                _lib.check(_lib.lib().sgs_fold_stats(
                    N, s_acc.data_ptr(), s_den.data_ptr(), max_radii.data_ptr(),
                    self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(), self.max_radii2D.data_ptr(),
                    torch.cuda.current_stream(bucket.device).cuda_stream), "sgs_fold_stats")
            else:
                self.xyz_gradient_accum += s_acc
                self.denom += s_den
                torch.maximum(self.max_radii2D, max_radii, out=self.max_radii2D)
                if reset_step:
                    s_acc.zero_(); s_den.zero_(); max_radii.zero_()
            return bucket[:n]

        return finish if async_op else finish()

    def sync_max(self):
        """With defer_max: the one all-reduce(MAX) that makes max_radii2D global; call before the
        statistics are read (gs_trainer.py:487-490 consumes them at densification time)."""
        if self.defer_max and self.world > 1:
            dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, group=self.group)
        return self.max_radii2D

    def reset(self):
        self.xyz_gradient_accum.zero_()
        self.denom.zero_()
        self.max_radii2D.zero_()


def broadcast_parameters(tensors: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Replicate the canonical Gaussians (and skinning weights) from rank `src`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the default group
    when WORLD_SIZE > 1 (backend nccl on CUDA, gloo otherwise)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # NCCL's CTA budget.  A synchronous step (the optimizer needs the reduced gradients before the
    # next forward) has the all-reduce on its critical path: it should be fast, so the cap is loose
    # (measured at 2 GPUs, frames/s of the synchronous step: cap 8 -> 2902, 16 -> 3688, 32 -> 4047;
    # profiles/README.md, round 2).  A pipelined exchange (overlapping the next frames) wants the
    # opposite -- thin, so it stays out of the way of the memory-bound head of the next frame: cap 8
    # (round 1: 8 GPUs 13548 frames/s at 32 CTAs, 14334 at 8).  SGS_NCCL_MAX_CTAS overrides; 0 = NCCL's own choice.
    cap = os.environ.get("SGS_NCCL_MAX_CTAS", "32")
    if cap != "0":
        os.environ.setdefault("NCCL_MAX_CTAS", cap)
    # the 52.8 MB gradient bucket at 8 GPUs (tools/nccl_ar_probe.sh, profiles/r02_nccl_allreduce_n8.txt):
    # NCCL's own choice 0.233 ms, Ring 0.206 ms, Tree 0.273 ms (NVLS is not offered on this box)
    os.environ.setdefault("NCCL_ALGO", "Ring")
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world

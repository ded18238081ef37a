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


def frame_to_uint8(img: torch.Tensor, out: Optional[torch.Tensor] = None, bgr: bool = False) -> torch.Tensor:
    """(3,H,W) float32 CUDA image -> (H,W,3) uint8 on the device (sgs_frame_to_u8): the reference's
    `(image.cpu().clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype('uint8')` (gs_trainer.py:716-717)
    and, with bgr=True, its `cv2.cvtColor(..., COLOR_RGB2BGR)` (:718) -- same bits, done BEFORE the
    device -> host copy, which then moves a quarter of the bytes."""
    from . import _lib
    if not img.is_cuda:
        raise _lib.SgsError("frame_to_uint8 needs a CUDA image (no CPU fallback)")
    img = img.detach()
    if img.dtype != torch.float32 or not img.is_contiguous():
        img = img.float().contiguous()
    _, H, W = img.shape
    if out is None:
        out = torch.empty(H, W, 3, device=img.device, dtype=torch.uint8)
    _lib.check(_lib.lib().sgs_frame_to_u8(img.data_ptr(), H, W, int(bgr), out.data_ptr(), _lib.raw_stream(img.device)),
               "sgs_frame_to_u8")
    return out


class FrameWriter:
    """Asynchronous output path of the animation loop (gs_trainer.py:716-728 writes every frame
    synchronously: blocking float32 D2H, numpy conversion, cv2.imwrite, all on the render thread).

    submit(image, name) converts the frame to 8 bits on the device, starts its copy into one of
    `depth` pinned host buffers on a side stream and returns at once; encoder threads pick the
    buffer up when its copy event has completed and write `<dir>/<name>.<ext>` with cv2 (the
    reference's encoder; BGR order is produced on the device).  Rendering, copy and encoding of
    different frames overlap; submit() only blocks when all `depth` buffers are in flight.
    `encode=False` stops after the copy (frames are handed to `sink(name, array)` if given)."""

    def __init__(self, out_dir: Optional[str], H: int, W: int, device, depth: int = 8, workers: int = 4,
                 ext: str = "jpg", encode: bool = True, sink=None):
        import queue
        import threading
        self.dir, self.ext, self.encode, self.sink = out_dir, ext, encode and out_dir is not None, sink
        if self.encode:
            import os
            os.makedirs(out_dir, exist_ok=True)
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.slots = [dict(dev=torch.empty(H, W, 3, device=self.dev, dtype=torch.uint8),
                           host=torch.empty(H, W, 3, dtype=torch.uint8).pin_memory(),
                           ready=torch.cuda.Event(), done=threading.Event()) for _ in range(depth)]
        for s in self.slots:
            s["done"].set()
        self.q = queue.Queue()
        self.n = 0
        self.bytes = 0
        self.errors = []
        self.threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        for th in self.threads:
            th.start()

    def _work(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            slot, name = item
            try:
                slot["ready"].synchronize()                    # this frame's copy has landed
                arr = slot["host"].numpy()
                if self.encode:
                    import cv2
                    if not cv2.imwrite(f"{self.dir}/{name}.{self.ext}", arr):
                        raise IOError(f"cv2.imwrite failed for {name}")
                elif self.sink is not None:
                    self.sink(name, arr)
            except Exception as e:                             # surfaced by close()
                self.errors.append(e)
            finally:
                slot["done"].set()

    def submit(self, image: torch.Tensor, name: str) -> None:
        slot = self.slots[self.n % len(self.slots)]
        self.n += 1
        slot["done"].wait()                                    # the encoder is finished with this buffer
        slot["done"].clear()
        frame_to_uint8(image, slot["dev"], bgr=self.encode)    # on the render stream, right behind the blend
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ev)
            slot["host"].copy_(slot["dev"], non_blocking=True)
            slot["ready"].record(self.copy_stream)
        self.bytes += slot["host"].numel()
        self.q.put((slot, name))

    def close(self) -> int:
        """Wait for every submitted frame; returns the number written.  Raises the first encoder error."""
        for _ in self.threads:
            self.q.put(None)
        for th in self.threads:
            th.join()
        if self.errors:
            raise self.errors[0]
        return self.n


def _render_lanes(steps, frames, lo: int, hi: int, out: torch.Tensor) -> None:
    """Frames lo..hi dealt round-robin to `steps`, each AvatarStep on its own stream: the per-Gaussian and
    binning kernels of one frame leave SMs idle that the other frame's kernels fill (measured with two
    frames in flight: 1.23x at 1080p / 200k Gaussians, bench.py key `concurrent`).  Every step owns its
    scratch, image buffer and pinned counters, so nothing crosses lanes; a block of frames is redone if a
    pair list overflowed in it."""
    from ._lib import SgsError
    dev, R = steps[0].dev, len(steps)
    cur = torch.cuda.current_stream(dev)
    lanes = [torch.cuda.Stream(dev) for _ in steps]
    start = torch.cuda.Event()
    start.record(cur)
    for ln in lanes:
        ln.wait_event(start)                 # `out` and the frames' tensors were produced on the caller's stream
    slots = min(s.COUNTER_SLOTS for s in steps)
    first = lo
    while first < hi:
        end = min(hi, first + slots * R)
        for f in range(first, end):
            k = (f - first) % R
            with torch.cuda.stream(lanes[k]):
                img = steps[k].forward(frames[f], slot=(f - first) // R)
                out[f - lo].copy_(img)
        for ln in lanes:
            ln.synchronize()
        grown = False
        for s in steps:
            try:
                s.check_capacity()
            except SgsError:
                grown = True                 # its capacity has been raised: redo this block
        if not grown:
            first = end
    for ln in lanes:                         # (already drained; keeps the caller's stream ordered after the lanes)
        cur.wait_stream(ln)


def render_frames(step, frames: Sequence[FrameInputs], rank: int = 0, world: int = 1,
                  out: Optional[torch.Tensor] = None, clamp: bool = True) -> Tuple[int, int, torch.Tensor]:
    """Render this rank's share of `frames`; returns (lo, hi, images (hi-lo, 3, H, W)).

    `out` (optional, (>= hi-lo, 3, H, W) on the device) receives the images; `clamp` applies
    the reference's `torch.clamp(rendered_image, 0, 1)` (gs_renderer_single.py:96).  The
    launch sequence per frame has no host synchronisation; the pair-list capacity is checked
    once at the end (and the affected frames re-rendered if it had to grow).
    `step` may be a list of AvatarSteps built over the same avatar: the frames are then dealt
    round-robin to them, each on its own stream (frames are independent; same images)."""
    if isinstance(step, (list, tuple)):
        steps = list(step)
        if len(steps) == 1:
            step = steps[0]
        else:
            if any((s.H, s.Wd, s.dev) != (steps[0].H, steps[0].Wd, steps[0].dev) for s in steps):
                raise ValueError("render_frames: the AvatarSteps must share image size and device")
            lo, hi = dp.shard_frames(len(frames), rank, world)
            n = hi - lo
            if out is None:
                out = torch.empty(max(n, 0), 3, steps[0].H, steps[0].Wd, device=steps[0].dev, dtype=torch.float32)
            keep = [s.forward_only for s in steps]
            for s in steps:
                s.forward_only = True
            try:
                _render_lanes(steps, frames, lo, hi, out)
            finally:
                for s, k in zip(steps, keep):
                    s.forward_only = k
            if clamp:
                out[:n].clamp_(0.0, 1.0)
            return lo, hi, out[:n]
    lo, hi = dp.shard_frames(len(frames), rank, world)
    n = hi - lo
    if out is None:
        out = torch.empty(max(n, 0), 3, step.H, step.Wd, device=step.dev, dtype=torch.float32)
    from ._lib import SgsError
    keep, step.forward_only = step.forward_only, True      # no backward follows an animation frame
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
    step.forward_only = keep
    if clamp:
        out[:n].clamp_(0.0, 1.0)
    return lo, hi, out[:n]

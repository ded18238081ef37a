"""AvatarStep -- the whole per-frame hot path (pose -> A -> LBS -> rasterize -> backward ->
LBS backward -> densification statistics) as one preallocated, sync-free launch sequence.

This is the fused fast path behind the two drop-in boundaries (`diff_gaussian_rasterization`
and `sings_b200.deform`): same kernels, same results, but without per-call tensor allocation
and autograd bookkeeping.  It corresponds to one iteration of the reference's hot loop between
`human_gs.forward` and `loss.backward()` (/root/reference/sings/rec/trainer/gs_trainer.py:
229-244, 400; sings_hybrid.py:398-428; gs_renderer_single.py:45-107) plus the statistics of
gs_trainer.py:486-492.

All canonical-parameter gradients land in ONE flat float32 bucket
    [ d_xyz_canon (N,3) | d_scales (N,3) | d_rotmat_canon (N,9) | d_opacity (N) | d_shs (N,M,3)
      | xyz_gradient_accum (N) | denom (N) ]
so the data-parallel exchange is a single SUM all-reduce of the bucket and a MAX all-reduce
of `max_radii2D` (sings_b200/dp.py).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .rasterizer import _sizes


@dataclass
class FrameInputs:
    """Per-frame inputs on the device (the reference's data dict)."""
    pose: torch.Tensor          # (J,3) axis-angle (joint 0 = global orientation)
    transl: torch.Tensor        # (3,)
    viewmatrix: torch.Tensor    # (4,4) W2C^T
    projmatrix: torch.Tensor    # (4,4) full projection
    campos: torch.Tensor        # (3,)
    bg: torch.Tensor            # (3,)
    tanfovx: float
    tanfovy: float
    smpl_scale: Optional[torch.Tensor] = None   # (1,)

    def __post_init__(self):
        # the library reads raw memory: every tensor must be float32 and C-contiguous.  (A camera
        # matrix built as `W2C.T` in numpy is an F-ordered view; torch.as_tensor keeps its strides and
        # data_ptr() would then hand the kernels the TRANSPOSE -- harmless only for an identity pose.)
        for name in ("pose", "transl", "viewmatrix", "projmatrix", "campos", "bg", "smpl_scale"):
            v = getattr(self, name)
            if v is not None and (v.dtype != torch.float32 or not v.is_contiguous()):
                setattr(self, name, v.float().contiguous())


class AvatarStep:
    def __init__(self, xyz_canon, rotmat_canon, scales, opacity, shs, lbs_weights, rest_joints,
                 parents, inv_A_t2cano, H: int, W: int, sh_degree: int, pair_capacity: int = 0,
                 timing: bool = False):
        dev = xyz_canon.device
        if dev.type != "cuda":
            raise _lib.SgsError("AvatarStep needs CUDA tensors (no CPU fallback)")
        self.L = _lib.lib()
        self.dev = dev
        f = lambda t: None if t is None else t.detach().float().contiguous()
        self.xyz_canon, self.rot_canon, self.scales = f(xyz_canon), f(rotmat_canon), f(scales)
        self.opacity, self.shs, self.W_lbs = f(opacity), f(shs), f(lbs_weights)
        self.rest, self.inv_A = f(rest_joints), f(inv_A_t2cano)
        self.parents = parents.to(device=dev, dtype=torch.int32).contiguous()
        self.N, self.J = self.xyz_canon.shape[0], self.W_lbs.shape[1]
        self.M = self.shs.shape[1]
        self.H, self.Wd, self.D = int(H), int(W), int(sh_degree)
        N, J, M = self.N, self.J, self.M
        e = lambda *s, dt=torch.float32: torch.empty(*s, device=dev, dtype=dt)
        # forward intermediates
        self.A = e(1, J, 4, 4)
        self.G = e(1, J, 12)
        self.xyz, self.rotq, self.sc = e(1, N, 3), e(1, N, 4), e(1, N, 3)
        self.color = e(3, H, W)
        self.radii = e(N, dt=torch.int32)
        self.L_cap = int(pair_capacity) if pair_capacity else max(8 * N, 1 << 16)
        self._alloc_scratch()
        # {num_rendered, overflow} of a forward, written by the device into mapped pinned memory.  One
        # row per in-flight frame (forward(slot=...)): a later forward must not overwrite the flag of
        # a frame that has not been examined yet (check_capacity looks at every row).
        self.COUNTER_SLOTS = 32
        self.counters = torch.zeros(self.COUNTER_SLOTS, 2, dtype=torch.int32).pin_memory()
        self._gen = 0                      # bumped when scratch is reallocated: captured graphs go stale
        # backward intermediates (rasterizer boundary gradients)
        self.g_means3D, self.g_means2D, self.g_colors = e(N, 3), e(N, 3), e(N, 3)
        self.g_cov = e(N, 6)
        self.g_scales_r, self.g_rots = e(N, 3), e(N, 4)
        # flat gradient bucket (see module docstring)
        self.n_rot = 9 if self.rot_canon is not None else 0
        sizes = [3 * N, 3 * N, self.n_rot * N, N, 3 * M * N, N, N]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        self.bucket = torch.zeros(int(offs[-1]), device=dev, dtype=torch.float32)
        v = lambda i, *shape: self.bucket[int(offs[i]):int(offs[i + 1])].view(*shape)
        self.d_xyz_canon, self.d_scales = v(0, N, 3), v(1, N, 3)
        self.d_rot_canon = v(2, N, 3, 3) if self.n_rot else None
        self.d_opacity, self.d_shs = v(3, N, 1), v(4, N, M, 3)
        self.grad_accum, self.denom = v(5, N), v(6, N)
        self.n_param_grads = int(offs[5])
        self.max_radii2D = torch.zeros(N, device=dev, dtype=torch.float32)
        # small per-frame gradients
        self.small_grads = torch.zeros(J * 16 + 4, device=dev)
        self.d_A = self.small_grads[:J * 16].view(1, J, 4, 4)
        self.d_transl = self.small_grads[J * 16:J * 16 + 3].view(1, 3)
        self.d_pose = e(1, J, 3)
        self._bwd_clean = False
        self._pack_weights()
        # shs is a model parameter: the kernels in front of the geometry kernels (LBS forward,
        # blend backward) are this library's and never write it, so its rows may be prefetched
        # ahead of the dependency wait (SGS_FLAG_EARLY_PARAMS)
        self._early = 0 if os.environ.get("SGS_NO_EARLY_PARAMS") else _lib.FLAG_EARLY_PARAMS
        # set for inference / animation (no backward() will follow a forward()): the forward then
        # skips the per-block lists and work items it otherwise leaves for the backward blend
        self.forward_only = False
        self._graphs = []
        self.timing = None
        # stage events are recorded only while this is set.  Inside a captured frame every
        # record is a graph node between two kernels, which turns their programmatic
        # (overlapped) launch edge into a full dependency: ~4 us per event, ~45 us per frame at
        # 200k Gaussians / 1024^2 (tools/probe_events.py) -- so capture(stages=False) for speed.
        self.record_stages = True
        if timing:
            h = C.c_void_p()
            _lib.check(self.L.sgs_timing_create(16, C.byref(h)), "sgs_timing_create")
            self.timing = h

    def _pack_weights(self):
        """Compact copy of lbs_weights for the fused kernels (include/sings_b200.h: packed skinning
        weights).  The buffer is constant between densifications (sings_hybrid.py:724); call
        again after it changes.  Rows with more than 16 non-zero joints (or SGS_NO_FUSE=1) keep the
        step on the unfused kernels."""
        L_, p = self.L, _lib.ptr
        self.K, self.wq, self.iq = 0, None, None
        if os.environ.get("SGS_NO_FUSE"):
            return
        st = torch.cuda.current_stream(self.dev).cuda_stream
        nnz = torch.zeros(1, device=self.dev, dtype=torch.int32)
        K = 4
        while K <= 16:
            nb = int(L_.sgs_lbs_packed_bytes(self.N, K))
            wq = torch.empty(nb // 4, device=self.dev, dtype=torch.float32)
            iq = torch.empty(max(nb // 16, 1), device=self.dev, dtype=torch.int32)
            nnz.zero_()
            _lib.check(L_.sgs_lbs_pack_weights(self.N, self.J, p(self.W_lbs), K, p(wq), p(iq), p(nnz), st),
                       "sgs_lbs_pack_weights")
            m = int(nnz.item())
            if m <= K:
                self.K, self.wq, self.iq = K, wq, iq
                return
            K = (m + 3) // 4 * 4

    def _deform_args(self, fr) -> "_lib.DeformArgs":
        p = _lib.ptr
        d = _lib.DeformArgs()
        d.N, d.J, d.K, d.rot6d = self.N, self.J, self.K, 0
        pose = fr.pose.reshape(1, self.J, 3)
        for name, tns in (("pose", pose), ("rest", self.rest), ("parents", self.parents),
                          ("inv_A_t2cano", self.inv_A), ("xyz_canon", self.xyz_canon), ("scales", self.scales),
                          ("rot_canon", self.rot_canon), ("wq", self.wq), ("iq", self.iq),
                          ("smpl_scale", fr.smpl_scale), ("transl", fr.transl), ("A", self.A), ("G", self.G),
                          ("xyz", self.xyz), ("rotq", self.rotq), ("scales_out", self.sc),
                          ("d_xyz_canon", self.d_xyz_canon), ("d_rot_canon", self.d_rot_canon),
                          ("d_scales", self.d_scales), ("d_A", self.d_A), ("d_transl", self.d_transl),
                          ("d_pose", self.d_pose)):
            setattr(d, name, p(tns))
        return d

    def _alloc_scratch(self):
        gb, bb, ib, ab = _sizes(self.N, self.Wd, self.H, self.L_cap)
        e = lambda n: torch.empty(n, device=self.dev, dtype=torch.uint8)
        self.geom, self.binning, self.img, self.acc = e(gb), e(bb), e(ib), e(ab)

    # ---------------------------------------------------------------------------------
    def forward(self, fr: FrameInputs, stream=None, slot: int = 0):
        """pose -> A -> LBS -> rasterize.  Returns the (3,H,W) image (a persistent buffer).
        `slot` selects the pinned {num_rendered, overflow} row this frame reports into."""
        L_, p = self.L, _lib.ptr
        cnt_ptr = self.counters.data_ptr() + 8 * (int(slot) % self.COUNTER_SLOTS)
        st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        self._fr = fr
        fo = _lib.FLAG_FORWARD_ONLY if self.forward_only else 0
        self._fwd_was_forward_only = bool(self.forward_only)
        pose = fr.pose.reshape(1, self.J, 3)
        tm = self.timing if self.record_stages else None
        # everything the frame needs zeroed is cleared here, up front: a memset between two
        # kernels would cost them their overlapped (programmatic dependent) launch
        _lib.check(L_.sgs_raster_clear(self.N, self.Wd, self.H, self.L_cap, p(self.binning), p(self.acc),
                                       p(self.small_grads), self.small_grads.numel() * 4, st),   # d_A and d_transl
                   "sgs_raster_clear")
        self._bwd_clean = True             # one backward may rely on the up-front clearing
        if self.K:
            # the fused path: pose -> A, then LBS inside the rasterizer's preprocess kernel
            d = self._deform_args(fr)
            _lib.check(L_.sgs_avatar_forward(
                C.byref(d), self.D, self.M, self.Wd, self.H, p(fr.bg), p(self.opacity), 1.0, p(fr.viewmatrix),
                p(fr.projmatrix), p(fr.campos), float(fr.tanfovx), float(fr.tanfovy), p(self.shs), self.L_cap,
                p(self.geom), p(self.binning), p(self.img), p(self.color), p(self.radii), None, None,
                cnt_ptr, st, _lib.FLAG_PRECLEARED | self._early | fo, tm), "sgs_avatar_forward")
            return self.color
        if tm:
            L_.sgs_timing_record(tm, 8, st)
        _lib.check(L_.sgs_pose_lbs_fwd(p(pose), p(self.rest), p(self.parents), p(self.inv_A), 1, self.N,
                                       self.J, p(self.A), p(self.G), p(self.xyz_canon), p(self.W_lbs),
                                       p(self.rot_canon), p(self.scales), p(fr.smpl_scale), p(fr.transl),
                                       p(self.xyz), p(self.rotq), p(self.sc), st), "sgs_pose_lbs_fwd")
        if tm:
            L_.sgs_timing_record(tm, 9, st)
        _lib.check(L_.sgs_raster_forward(
            self.N, self.D, self.M, self.Wd, self.H, p(fr.bg), p(self.xyz), None, p(self.opacity),
            p(self.sc), 1.0, p(self.rotq), None, p(fr.viewmatrix), p(fr.projmatrix), p(fr.campos),
            float(fr.tanfovx), float(fr.tanfovy), p(self.shs), 0, self.L_cap, p(self.geom),
            p(self.binning), p(self.img), p(self.color), p(self.radii), None, None,
            cnt_ptr, st, _lib.FLAG_PRECLEARED | self._early | fo, tm), "sgs_raster_forward")
        return self.color

    def backward(self, dL_dimage: torch.Tensor, stream=None, stats: bool = True):
        """Backward of forward() for dL/d(image) (3,H,W); fills the gradient bucket, d_pose,
        d_transl and (stats=True) accumulates the densification statistics of this view."""
        L_, p = self.L, _lib.ptr
        st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        fr = self._fr
        if getattr(self, "_fwd_was_forward_only", False):
            raise _lib.SgsError("backward() after a forward_only forward(): the forward left no block lists for it")
        tm = self.timing if self.record_stages else None
        flags = (_lib.FLAG_PRECLEARED if self._bwd_clean else 0) | self._early
        if not self._bwd_clean:            # a second backward of the same forward: clear again
            self.small_grads.zero_()
        self._bwd_clean = False
        if self.K:
            d = self._deform_args(fr)
            _lib.check(L_.sgs_avatar_backward(
                C.byref(d), self.D, self.M, self.Wd, self.H, p(fr.bg), 1.0, p(fr.viewmatrix), p(fr.projmatrix),
                p(fr.campos), float(fr.tanfovx), float(fr.tanfovy), p(self.shs), p(self.radii), p(dL_dimage),
                self.L_cap, p(self.geom), p(self.binning), p(self.img), p(self.acc), p(self.g_means2D),
                p(self.d_opacity), p(self.d_shs), p(self.grad_accum) if stats else None,
                p(self.denom) if stats else None, p(self.max_radii2D) if stats else None, st, flags, tm),
                "sgs_avatar_backward")
            return
        _lib.check(L_.sgs_raster_backward(
            self.N, self.D, self.M, self.Wd, self.H, p(fr.bg), p(self.xyz), None, p(self.sc), 1.0,
            p(self.rotq), None, p(fr.viewmatrix), p(fr.projmatrix), p(fr.campos), float(fr.tanfovx),
            float(fr.tanfovy), p(self.shs), p(self.radii), p(dL_dimage), self.L_cap, p(self.geom),
            p(self.binning), p(self.img), p(self.acc), p(self.g_means3D), p(self.g_means2D),
            p(self.g_colors), p(self.d_opacity), p(self.g_cov), p(self.d_shs), p(self.g_scales_r),
            p(self.g_rots), p(self.grad_accum) if stats else None, p(self.denom) if stats else None,
            p(self.max_radii2D) if stats else None, st, flags, tm), "sgs_raster_backward")
        if tm:
            L_.sgs_timing_record(tm, 10, st)
        pose = fr.pose.reshape(1, self.J, 3)
        _lib.check(L_.sgs_lbs_bwd(1, self.N, self.J, p(self.A), p(self.xyz_canon), p(self.W_lbs),
                                  p(self.rot_canon), p(self.scales), p(fr.smpl_scale), p(fr.transl),
                                  None, None, None, p(self.g_means3D), p(self.g_rots),
                                  p(self.g_scales_r), None, p(self.d_xyz_canon), p(self.d_rot_canon),
                                  p(self.d_scales), p(self.d_A), None, p(self.d_transl), st),
                   "sgs_lbs_bwd")
        _lib.check(L_.sgs_pose_to_A_bwd(p(pose), p(self.rest), p(self.parents), p(self.inv_A),
                                        p(self.G), p(self.d_A), 1, self.J, p(self.d_pose), st),
                   "sgs_pose_to_A_bwd")
        if tm:
            L_.sgs_timing_record(tm, 11, st)

    def capture(self, fr: FrameInputs, dL_dimage: Optional[torch.Tensor], loss_weight: Optional[torch.Tensor] = None,
                prologue=None, stages: bool = True, stage_mask: Optional[int] = None, loss_fn=None):
        """Record forward(fr) [+ loss = <image, loss_weight>] + backward(dL_dimage) into a CUDA
        graph and return replay() (dL_dimage None: the forward alone, e.g. animation frames).  The launch sequence is static -- capacity-sized pair list,
        device-side counts, no host round trip -- so the whole frame becomes one graph launch.
        `fr`'s tensors and `dL_dimage` are captured by address: refresh their contents in
        place before each replay.  Stage events keep working (external event-record nodes).
        `prologue()` (optional) runs first inside the graph -- e.g. the caller's decoding of an
        uploaded target image into `dL_dimage` / `loss_weight`.  stages=False leaves the stage
        event records out of the graph (see record_stages).
        `loss_fn(image) -> (loss, dL_dimage)` (optional, instead of dL_dimage / loss_weight) runs
        between forward and backward inside the graph -- e.g. sings_b200.losses.ImageLossBuffers."""
        keep, self.record_stages = self.record_stages, bool(stages)
        if self.timing and stage_mask is not None:     # record only these stage events in this graph
            self.L.sgs_timing_set_mask(self.timing, int(stage_mask) & 0xffffffff)
        try:
            return self._capture(fr, dL_dimage, loss_weight, prologue, loss_fn)
        finally:
            self.record_stages = keep
            if self.timing and stage_mask is not None:
                self.L.sgs_timing_set_mask(self.timing, 0xffffffff)

    def _capture(self, fr, dL_dimage, loss_weight, prologue, loss_fn=None):
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside the capture (lazy initialisation)
            if prologue is not None:
                prologue()
            img = self.forward(fr)
            if loss_fn is not None:
                self.loss, dL_dimage = loss_fn(img)
            if loss_weight is not None:
                self.loss = torch.dot(img.view(-1), loss_weight.view(-1))
            if dL_dimage is not None:
                self.backward(dL_dimage)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        if prologue is None and loss_weight is None and loss_fn is None and not os.environ.get("SGS_TORCH_GRAPH"):
            # the pure C-ABI frame: recorded through the library (no torch op inside), replayed
            # with a single cudaGraphLaunch on the current stream
            h = C.c_void_p()
            _lib.check(self.L.sgs_graph_begin(side.cuda_stream), "sgs_graph_begin")
            try:
                self.forward(fr, stream=side)
                if dL_dimage is not None:
                    self.backward(dL_dimage, stream=side)
            finally:
                rc = self.L.sgs_graph_end(side.cuda_stream, C.byref(h))
            _lib.check(rc, "sgs_graph_end")
            self._graphs.append(h)
            launch, dev, gen = self.L.sgs_graph_launch, self.dev, self._gen

            def replay():
                self._check_gen(gen)
                _lib.check(launch(h, torch.cuda.current_stream(dev).cuda_stream), "sgs_graph_launch")
            return replay
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if prologue is not None:
                prologue()
            img = self.forward(fr)
            if loss_fn is not None:
                self.loss, dL_dimage = loss_fn(img)
            if loss_weight is not None:
                self.loss = torch.dot(img.view(-1), loss_weight.view(-1))
            if dL_dimage is not None:
                self.backward(dL_dimage)
        self.graph = g
        gen = self._gen

        def replay_torch():
            self._check_gen(gen)
            g.replay()
        return replay_torch

    def _check_gen(self, gen: int) -> None:
        if gen != self._gen:
            raise _lib.SgsError("this captured frame is stale: the scratch buffers it was recorded with were "
                                "reallocated (pair-list capacity raised); call capture() again")

    def reset_stats(self):
        self.grad_accum.zero_()
        self.denom.zero_()
        self.max_radii2D.zero_()

    def check_capacity(self) -> int:
        """After a synchronisation: (num_rendered); grows the pair list and raises if the last
        forward overflowed (the frame must then be re-rendered)."""
        L, ovf = int(self.counters[:, 0].max()), bool(self.counters[:, 1].any())
        self.counters.zero_()
        if ovf:
            self.L_cap = int(L * 1.3) + 4096
            # graphs recorded with the old buffers must not be replayed: their pointers and L_cap are baked in
            for h in self._graphs:
                self.L.sgs_graph_destroy(h)
            self._graphs = []
            self.graph = None
            self._gen += 1
            self._alloc_scratch()
            raise _lib.SgsError(f"pair list overflowed (needed {L}); capacity raised to {self.L_cap}: re-render the "
                                f"frame(s) and capture() again")
        return L

    def image_state(self) -> dict:
        """final_T (H,W) float32 and n_contrib (H,W) int32 of the last forward ([upstream] imgBuffer contents)."""
        from .rasterizer import layout_info
        info = layout_info(self.N, self.Wd, self.H, self.L_cap)
        n = self.H * self.Wd
        raw = self.img
        return {"final_T": raw[info["final_T"]:info["final_T"] + 4 * n].view(torch.float32).view(self.H, self.Wd).clone(),
                "n_contrib": raw[info["n_contrib"]:info["n_contrib"] + 4 * n].view(torch.int32).view(self.H, self.Wd).clone()}

    def interval_ms(self, i: int, j: int) -> float:
        """Device time between stage events i and j of the last frame that recorded both."""
        ms = C.c_float()
        _lib.check(self.L.sgs_timing_elapsed_ms(self.timing, i, j, C.byref(ms)), "elapsed")
        return float(ms.value)

    def launches_per_frame(self, backward: bool = True) -> int:
        """Kernels of this library that one frame launches (the bench line's gpu_launches)."""
        info = (C.c_longlong * 16)()
        _lib.check(self.L.sgs_raster_layout_info(self.N, self.Wd, self.H, self.L_cap, info), "layout")
        tile_passes = int(info[10]) - 4
        fwd = 1 + 1 + (0 if self.K else 1) + 1 + 4 + 1 + tile_passes + 1 + 1   # clear, pose->A, [LBS], geometry, 4 depth passes, emit, tile passes, ranges, blend
        bwd = 1 + 1 + (0 if self.K else 1) + 1                                   # blend bwd, geometry bwd, [LBS bwd], pose bwd
        return fwd + (bwd if backward else 0)

    def stage_ms(self, backward: bool = True) -> dict:
        """Per-stage device times of the last forward [+ backward] (needs timing=True)."""
        if not self.timing:
            raise _lib.SgsError("AvatarStep(timing=True) required")
        out = {}
        ms = C.c_float()
        if self.K:      # fused kernels: deform + preprocess is one stage, so is their backward
            stages = [("deform_geometry", 8, 1), ("binning", 1, 3), ("blend_fwd", 3, 4), ("blend_bwd", 5, 6),
                      ("geometry_lbs_bwd", 6, 11), ("total", 8, 11)]
        else:
            stages = [("lbs_fwd", 8, 9), ("geometry", 0, 1), ("binning", 1, 3), ("blend_fwd", 3, 4),
                      ("blend_bwd", 5, 6), ("geometry_bwd", 6, 7), ("lbs_bwd", 10, 11), ("total", 8, 11)]
        if not backward:     # forward-only frame (animation): the stages up to the blend
            stages = [s for s in stages if s[0] in ("deform_geometry", "lbs_fwd", "geometry", "binning", "blend_fwd")]
            stages.append(("total", 8, 4))
        for name, i, j in stages:
            _lib.check(self.L.sgs_timing_elapsed_ms(self.timing, i, j, C.byref(ms)), "elapsed")
            out[name] = float(ms.value)
        return out

    def __del__(self):
        try:
            for h in self._graphs:
                self.L.sgs_graph_destroy(h)
        except Exception:
            pass
        try:
            if self.timing:
                self.L.sgs_timing_destroy(self.timing)
        except Exception:
            pass

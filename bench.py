#!/usr/bin/env python
"""bench.py -- deformed-avatar fwd+bwd frames/s (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W               # this repo's CUDA path
    torchrun ... bench.py --gpus N --steps K --warmup W         # one rank per GPU, views sharded
    python bench.py --impl reference --steps K --warmup W       # CPU arm (oracle port)

A step = one pass of the hot path over one synthetic frame: pose -> A -> LBS of every
Gaussian -> rasterize (1024^2, SH degree 3) -> dL/dimage -> rasterizer backward -> LBS
backward -> densification statistics (+ the gradient all-reduce when N > 1).
Workload = BASELINE.json configs[1]: 200k Gaussians, J=24, random pose, synthetic data.

`value`  : device-resident inputs, AvatarStep (sync-free launch sequence), CUDA events.
`e2e`    : the public API (sings_b200.deform + diff_gaussian_rasterization autograd) with HOST
           buffers: per step pose/transl/dL-dimage are copied from pinned memory and the loss
           is read back, all inside the timed region.
L2      : inputs rotate over a ring of distinct avatars larger than the 126 MB L2.
Nothing here reads /root/reference.  The CPU arm / cpu_baseline time the oracle port
(oracle/lbs_oracle.py + oracle/c/raster_oracle.c, OpenMP) -- the only use of oracle/ here.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GAUSS, H_IMG, W_IMG, SH_DEG, N_JOINTS = 200_000, 1024, 1024, 3, 24
RING = 4
U8_SCALE = 42.0       # uint8 target = rint(G * 42 + 127.5): +-3 sigma of the N(0,1) gradient image in 8 bits
WORKLOAD = "single B200: LBS + rasterize fwd+bwd, 200k Gaussians, SH degree 3, 1024x1024 view, random pose"
METRIC = "deformed-avatar fwd+bwd frames/s @200k Gaussians 1024^2"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_frame_inputs(seed: int, n=N_GAUSS, H=H_IMG, W=W_IMG, J=N_JOINTS):
    """numpy inputs of one avatar + one frame."""
    from sings_b200 import synthetic as syn
    av = syn.make_avatar(n, J, seed=seed)
    pose = syn.random_pose(J, seed=seed + 2)
    transl = syn.default_transl(H)
    view = syn.make_view(H, W)
    rng = np.random.default_rng(seed + 3)
    G = rng.normal(size=(3, H, W)).astype(np.float32)
    bg = np.ones(3, np.float32)
    return av, pose, transl, view, G, bg


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference torch LBS restated + C rasterizer), all host threads
# ------------------------------------------------------------------------------------------
def cpu_frame(av, pose, transl, view, G, bg, D=SH_DEG):
    import torch
    from oracle import lbs_oracle as lo
    from oracle import raster_oracle as ro
    t = torch.from_numpy
    xyz_c = t(av.xyz_canon).requires_grad_(True)
    sc_c = t(av.scales).requires_grad_(True)
    rot_c = t(av.rotmat_canon).requires_grad_(True)
    pose_t = t(pose)[None].clone().requires_grad_(True)
    tr = t(transl)[None].clone().requires_grad_(True)
    A = lo.pose_to_A(pose_t, t(av.rest), av.parents, t(av.inv_A_t2cano))
    xyz, q, sc, _ = lo.deform(A, xyz_c, t(av.lbs_weights), sc_c, rot_c, None, tr)
    cam = ro.Camera(W=view.image_width, H=view.image_height, tanfovx=view.tanfovx, tanfovy=view.tanfovy,
                    view=view.world_view_transform.reshape(-1), proj=view.full_proj_transform.reshape(-1),
                    campos=view.camera_center)
    st = ro.forward(cam, xyz[0].detach().numpy(), av.opacity, bg, shs=av.shs,
                    scales=sc[0].detach().numpy(), rotations=q[0].detach().numpy(), sh_degree=D)
    gr = ro.backward(st, G)
    torch.autograd.backward([xyz, q, sc], [t(gr["means3D"])[None], t(gr["rotations"])[None], t(gr["scales"])[None]])
    return st.num_rendered, float((st.color * G).sum())


def run_cpu(steps: int, warmup: int, frames):
    import torch
    from oracle import raster_oracle as ro
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ro.set_threads(cores)
    for i in range(warmup):
        cpu_frame(*frames[i % len(frames)])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_frame(*frames[i % len(frames)])
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = [build_frame_inputs(100)]
    steps, warmup = args.steps, args.warmup
    # bound the run to a few minutes: probe one frame, then cap the step count
    t0 = time.perf_counter()
    cpu_frame(*frames[0])
    t1 = time.perf_counter() - t0
    budget = 150.0
    steps_run = max(1, min(steps, int(budget / max(t1, 1e-3))))
    warm_run = min(warmup, 1)
    fps, ms, cores = run_cpu(steps_run, warm_run, frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps_run, "warmup": warm_run, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "image": [H_IMG, W_IMG],
                   "sh_degree": SH_DEG, "joints": N_JOINTS},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{steps_run} full frames (LBS fwd+bwd via the torch restatement of the "
                                   f"reference's lbs_extra/rotations, rasterizer fwd+bwd via the OpenMP C oracle); "
                                   f"the reference has no CPU rasterizer and its CUDA one is not vendored"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def algorithmic_bytes(N, J, M, L, W, H, n_vis, passes=6):
    """Per-launch algorithmic bytes of each stage (DESIGN.md 'Algorithmic bytes')."""
    pix, tiles = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    return {
        "lbs_fwd": N * (100 + 4 * J),
        "lbs_bwd": N * (148 + 4 * J),
        "geometry": N * (236 + 75),                                 # preprocessCUDA
        # InclusiveSum + duplicateWithKeys + SortPairs + identifyTileRanges: one stage
        # here, because the count / scan / scatter binning produces list and ranges together
        "binning": 8 * N + 20 * N + 12 * L + (8 + 24 * passes) * L + 8 * L + 8 * tiles,
        "blend_fwd": 40 * L + 20 * pix + 8 * tiles,
        "blend_bwd": 40 * L + 20 * pix + 36 * n_vis,
        "geometry_bwd": (300 + 256) * n_vis,
    }


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from sings_b200 import dp
    from sings_b200.step import AvatarStep, FrameInputs

    rank, local, world = dp.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hbm_peak, peak_src = peaks()

    # ---- workload: ring of distinct avatars (inputs larger than L2), each rank its own views
    sets = []
    for r in range(RING):
        av, pose, transl, view, G, bg = build_frame_inputs(1000 * rank + 10 * r)
        t = lambda a: torch.as_tensor(a, device=dev)
        step = AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs),
                          t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano),
                          H_IMG, W_IMG, SH_DEG, timing=True)
        fr = FrameInputs(pose=t(pose), transl=t(transl), viewmatrix=t(view.world_view_transform),
                         projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(bg),
                         tanfovx=view.tanfovx, tanfovy=view.tanfovy)
        sets.append(dict(step=step, fr=fr, G=t(G), av=av, pose=pose, transl=transl, view=view, G_np=G, bg=bg))
    exch = dp.GradExchange(N_GAUSS, sets[0]["step"].n_param_grads, dev,
                           defer_max=not os.environ.get("SGS_DP_MAX_EVERY_STEP")) if world > 1 else None

    def one_step(i, pending, staged=False):
        s = sets[i % RING]
        st = s["step"]
        st.record_stages = bool(staged)    # eager launches: stage events only in the instrumented passes
        if exch is not None and pending[i % RING] is not None:
            pending[i % RING]()            # finish the all-reduce that last used this bucket (folds + clears the step statistics)
            pending[i % RING] = None
        if "replay" in s:
            s[{False: "replay", True: "replay_staged", "coarse": "replay_coarse"}[staged]]()   # the frame as one CUDA-graph launch
        else:
            st.forward(s["fr"])
            st.backward(s["G"])
        if exch is not None:
            pending[i % RING] = exch.exchange(st.bucket, st.max_radii2D, async_op=True, reset_step=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pending = [None] * RING
    from sings_b200._lib import SgsError
    for attempt in range(4):
        for i in range(max(args.warmup, 3, RING)):
            one_step(i, pending)
        torch.cuda.synchronize()
        try:
            for s in sets:
                s["L"] = s["step"].check_capacity()
            break
        except SgsError:          # pair list capacity grown; warm up again with the larger buffers
            for s in sets:
                try:
                    s["step"].check_capacity()
                except SgsError:
                    pass
    if not args.no_graph:
        # capacities are settled: record each avatar's frame (forward + backward) as a CUDA graph
        for i in range(RING):
            if pending[i] is not None:
                pending[i]()
                pending[i] = None
        for s in sets:
            s["replay"] = s["step"].capture(s["fr"], s["G"], stages=False)
            s["replay_staged"] = s["step"].capture(s["fr"], s["G"], stages=True)
            # coarse: only the events around LBS + preprocess + sort + ranges (8 = frame start, 3 = after ranges)
            s["replay_coarse"] = s["step"].capture(s["fr"], s["G"], stages=True, stage_mask=(1 << 8) | (1 << 3))
        for i in range(2 * RING):
            one_step(i, pending)
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step(i, pending)
    for i in range(RING):
        if pending[i] is not None:
            pending[i]()
            pending[i] = None
    if exch is not None:
        exch.sync_max()                    # the deferred all-reduce(MAX) of max_radii2D, inside the timed region
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    for s in sets:
        s["step"].check_capacity()
    value = world * args.steps / (ms_total / 1e3)

    # ---- per-stage device times: a second timed pass over the same frames with CUDA events
    # recorded at the stage boundaries.  The records sit between kernels (graph nodes), which
    # turns the kernels' overlapped programmatic launch into full dependencies -- ~4 us per
    # event, ~45 us per frame (tools/probe_events.py) -- so they stay out of the headline region
    # above; the stage times below therefore sum to more than ms_per_step.
    n_stage = max(RING, min(args.steps, 100))
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(n_stage):
        one_step(i, pending, staged=True)
    for i in range(RING):
        if pending[i] is not None:
            pending[i]()
            pending[i] = None
    s1.record()
    barrier()
    ms_staged = s0.elapsed_time(s1) / n_stage
    stage = {}
    for s in sets:
        for k, v in s["step"].stage_ms().items():
            stage.setdefault(k, []).append(v)
    # third pass: one interval over the north star's target set (LBS + preprocess + sort + ranges),
    # two event records instead of five inside it -- less perturbation than the sum of its stages
    hot_coarse = None
    if not args.no_graph:
        for i in range(n_stage):
            one_step(i, pending, staged="coarse")
        for i in range(RING):
            if pending[i] is not None:
                pending[i]()
                pending[i] = None
        barrier()
        hot_coarse = float(np.mean([s["step"].interval_ms(8, 3) for s in sets]))
    stage = {k: float(np.mean(v)) for k, v in stage.items()}
    Lm = float(np.mean([s["L"] for s in sets]))
    n_vis = float(np.mean([int((s["step"].radii > 0).sum().item()) for s in sets]))
    ab = algorithmic_bytes(N_GAUSS, N_JOINTS, 16, Lm, W_IMG, H_IMG, n_vis)
    if "deform_geometry" in stage:      # fused kernels: the stages merge, their algorithmic bytes add
        ab["deform_geometry"] = ab.pop("lbs_fwd") + ab.pop("geometry")
        ab["geometry_lbs_bwd"] = ab.pop("geometry_bwd") + ab.pop("lbs_bwd")
    stages_out = {}
    for k, b in ab.items():
        ms = stage.get(k)
        if ms and ms > 0:
            gbs = b / (ms * 1e-3) / 1e9
            stages_out[k] = {"ms": round(ms, 4), "alg_mb": round(b / 1e6, 2), "gbs": round(gbs, 1),
                             "frac": round(gbs / hbm_peak, 4)}
    dom = max((k for k in stages_out), key=lambda k: stages_out[k]["ms"])
    hot = [k for k in ("lbs_fwd", "geometry", "deform_geometry", "binning") if k in ab]
    hot_b = sum(ab[k] for k in hot)
    hot_ms = sum(stage[k] for k in hot)
    hot_one = hot_coarse if hot_coarse else hot_ms
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):      # DRAM bytes per launch from the committed `ncu --set full` capture
        with open(tp) as f:
            tj = json.load(f)
        kern = {"deform_geometry": "geometry_kernel", "geometry_lbs_bwd": "geometry_bwd_kernel",
                "lbs_fwd": "lbs_fwd_kernel", "lbs_bwd": "lbs_bwd_kernel", "geometry": "geometry_kernel",
                "binning": "bin_scatter_kernel", "blend_fwd": "blend_fwd_kernel",
                "blend_bwd": "blend_bwd_kernel", "geometry_bwd": "geometry_bwd_kernel"}[dom]
        if kern in tj.get("kernels", {}):
            traffic, traffic_src = tj["kernels"][kern]["dram_bytes"], f"profiles/{tj.get('source')} ({kern})"
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": stages_out[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
        "frac": stages_out[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes": int(ab[dom]), "peak_source": peak_src,
        "stages": stages_out,
        "lbs_preprocess_sort": {"alg_mb": round(hot_b / 1e6, 2), "ms": round(hot_one, 4),
                                "gbs": round(hot_b / (hot_one * 1e-3) / 1e9, 1),
                                "frac": round(hot_b / (hot_one * 1e-3) / 1e9 / hbm_peak, 4),
                                "ms_sum_of_stages": round(hot_ms, 4),
                                "timing": "one interval, frame start -> after tile ranges (2 event records)"
                                          if hot_coarse else "sum of the four stage intervals"},
        "frame_alg_mb": round(sum(ab.values()) / 1e6, 1),
        "pairs_L": Lm, "visible": n_vis,
        "stage_timing": {"steps": n_stage, "ms_per_step": round(ms_staged, 4),
                         "note": "stage times come from a second timed pass of the same frames with CUDA events at the "
                                 "stage boundaries; the event records break the kernels' programmatic dependent launch "
                                 "(~4 us each), so that pass is slower than the headline region, which has none"},
    }

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, sets, dev, world, rank, exch)
    clocks = sampler.stop() if rank == 0 else None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        s = sets[0]
        frames = [(s["av"], s["pose"], s["transl"], s["view"], s["G_np"], s["bg"])]
        # bounded sample of the same workload: ~15 s of host work (probe one frame, then size the run)
        t0 = time.perf_counter()
        cpu_frame(*frames[0])
        n_cpu = max(2, min(200, int(15.0 / max(time.perf_counter() - t0, 1e-3))))
        fps, ms, cores = run_cpu(n_cpu, 0, frames)
        cpu_base = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                    "sample": f"{n_cpu} full frames of the same workload on the host (torch restatement of the "
                              f"reference LBS + OpenMP C rasterizer oracle, fwd+bwd)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "image": [H_IMG, W_IMG], "sh_degree": SH_DEG,
                       "joints": N_JOINTS, "views_per_gpu_per_step": 1,
                       "launch": "eager launches" if args.no_graph else "one CUDA graph per frame (forward+backward)",
                       "l2": f"inputs rotate over a ring of {RING} distinct avatars (~{RING * 70} MB of inputs) > 126 MB L2",
                       "parallelism": f"dp{world} (views sharded, gradient bucket all-reduced)" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
            "gpu_launches": 16 * args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, sets, dev, world, rank, exch=None):
    """Same metric end to end with HOST buffers.  Per step, inside the timed region: H2D of
    pose + transl + dL/dimage from pinned memory, forward, loss, backward, D2H of the loss.

    Two callers of the same kernels are timed:
      * `e2e` (headline): the C-ABI path -- `AvatarStep` drives include/sings_b200.h directly
        with preallocated buffers (what a trainer integrating the library calls per frame);
      * `e2e.dropin`: the reference-facing autograd modules (`sings_b200.deform` +
        `diff_gaussian_rasterization.GaussianRasterizer`), which add torch.autograd and
        per-call allocation overhead on the host.
    Both use the same input pipeline: inputs are prefetched one step ahead on a copy stream
    into double-buffered device staging (like a data loader with pinned memory) and the loss
    is read back through pinned slots one step late (like a logging trainer)."""
    import torch
    import torch.distributed as dist
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from sings_b200 import deform
    from sings_b200 import rasterizer as R
    from sings_b200.step import FrameInputs

    for s in sets:
        s["step"].record_stages = False     # no stage events on this path (eager launches included)
    host = []
    for s in sets:
        av = s["av"]
        t = lambda a: torch.as_tensor(a, device=dev)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        # the same dense gradient image quantised to 8 bits: how a target image is stored on disk
        t8 = np.clip(np.rint(s["G_np"] * U8_SCALE + 127.5), 0, 255).astype(np.uint8)
        host.append(dict(pose=pin(s["pose"]), transl=pin(s["transl"]), G=pin(s["G_np"]), T8=pin(t8), view=s["view"],
                         bg=t(s["bg"]), vm=t(s["view"].world_view_transform),
                         pm=t(s["view"].full_proj_transform), cp=t(s["view"].camera_center)))
    h2d = int(host[0]["pose"].numel() * 4 + host[0]["transl"].numel() * 4 + host[0]["G"].numel() * 4)
    cur = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    NBUF = 2
    stage = [dict(pose=torch.empty(N_JOINTS, 3, device=dev), transl=torch.empty(3, device=dev),
                  G=torch.empty(3, H_IMG, W_IMG, device=dev),
                  T8=torch.empty(3, H_IMG, W_IMG, device=dev, dtype=torch.uint8), ready=torch.cuda.Event(),
                  free=torch.cuda.Event()) for _ in range(NBUF)]
    loss_host = torch.zeros(NBUF).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(NBUF)]
    losses = []
    # diagnostics only (tools/e2e_probe.sh): "nog" skips the dL/dimage upload, "nosync" the lagged loss read
    PROBE = os.environ.get("SGS_E2E_PROBE", "")
    upload = {"u8": False}      # True: the per-step image upload is uint8 and decoded on the device

    def prefetch(i):
        hs, sb = host[i % RING], stage[i % NBUF]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sb["free"])          # the step that last used this buffer is done
            sb["pose"].copy_(hs["pose"], non_blocking=True)
            sb["transl"].copy_(hs["transl"], non_blocking=True)
            if upload["u8"]:
                sb["T8"].copy_(hs["T8"], non_blocking=True)
            elif "nog" not in PROBE:
                sb["G"].copy_(hs["G"], non_blocking=True)
            sb["ready"].record(copy_stream)

    def finish_step(i, loss, sb):
        sb["free"].record(cur)
        loss_host[i % NBUF:i % NBUF + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ev[i % NBUF].record(cur)
        if i >= 1 and "nosync" not in PROBE:           # read the previous step's loss
            loss_ev[(i - 1) % NBUF].synchronize()
            losses.append(float(loss_host[(i - 1) % NBUF]))

    # ---- C-ABI path ----
    pending = [None] * RING

    def drain_exchange():
        for k in range(RING):
            if pending[k] is not None:
                pending[k]()
                pending[k] = None

    assert RING % NBUF == 0        # ring slot k always meets staging buffer k % NBUF (graphs bind addresses)
    frames = [FrameInputs(pose=stage[k % NBUF]["pose"], transl=stage[k % NBUF]["transl"], viewmatrix=host[k]["vm"],
                          projmatrix=host[k]["pm"], campos=host[k]["cp"], bg=host[k]["bg"],
                          tanfovx=host[k]["view"].tanfovx, tanfovy=host[k]["view"].tanfovy) for k in range(RING)]
    replays = [None] * RING

    def step_abi(i, n_total):
        hs, sb, st = host[i % RING], stage[i % NBUF], sets[i % RING]["step"]
        if i + 1 < n_total:
            prefetch(i + 1)
        if exch is not None and pending[i % RING] is not None:
            pending[i % RING]()            # finish the all-reduce that last used this bucket (folds + clears the step statistics)
            pending[i % RING] = None
        cur.wait_event(sb["ready"])
        if replays[i % RING] is not None:
            replays[i % RING]()            # forward + loss + backward as one CUDA-graph launch
            loss = st.loss
        else:
            img = st.forward(frames[i % RING])
            loss = torch.dot(img.view(-1), sb["G"].view(-1))      # L = sum(image * G); dL/dimage = G
            st.backward(sb["G"])
        if exch is not None:
            pending[i % RING] = exch.exchange(st.bucket, st.max_radii2D, async_op=True, reset_step=True)
        finish_step(i, loss, sb)

    # ---- drop-in autograd path ----
    params = []
    if not args.no_dropin:
        for s in sets:
            av = s["av"]
            t = lambda a: torch.as_tensor(a, device=dev)
            params.append(dict(xyz=t(av.xyz_canon).requires_grad_(True), rot=t(av.rotmat_canon).requires_grad_(True),
                               scales=t(av.scales).requires_grad_(True), opacity=t(av.opacity).requires_grad_(True),
                               shs=t(av.shs).requires_grad_(True), W=t(av.lbs_weights), rest=t(av.rest),
                               parents=torch.from_numpy(av.parents).to(device=dev, dtype=torch.int32),
                               inv_A=t(av.inv_A_t2cano)))

    def step_dropin(i, n_total):
        hs, sb, p = host[i % RING], stage[i % NBUF], params[i % RING]
        if i + 1 < n_total:
            prefetch(i + 1)
        cur.wait_event(sb["ready"])
        pose = sb["pose"].detach().requires_grad_(True)
        transl = sb["transl"].detach().requires_grad_(True)
        A = deform.pose_to_A(pose, p["rest"], p["parents"], p["inv_A"])
        xyz, rotq, sc = deform.deform_gaussians(A, p["xyz"], p["W"], p["rot"], p["scales"], None, transl)
        v = hs["view"]
        rs = GaussianRasterizationSettings(
            image_height=v.image_height, image_width=v.image_width, tanfovx=v.tanfovx, tanfovy=v.tanfovy,
            bg=hs["bg"], scale_modifier=1.0, viewmatrix=hs["vm"], projmatrix=hs["pm"], sh_degree=SH_DEG,
            campos=hs["cp"], prefiltered=False, debug=False)
        means2D = torch.zeros_like(xyz, requires_grad=True)
        img, radii = GaussianRasterizer(rs)(means3D=xyz, means2D=means2D, shs=p["shs"], opacities=p["opacity"],
                                            scales=sc, rotations=rotq)
        loss = (img * sb["G"]).sum()
        for q in (p["xyz"], p["rot"], p["scales"], p["opacity"], p["shs"]):
            q.grad = None
        loss.backward()
        finish_step(i, loss, sb)

    def run(step, n):
        for sb in stage:
            sb["free"].record(cur)
        prefetch(0)
        for i in range(n):
            step(i, n)
        loss_ev[(n - 1) % NBUF].synchronize()
        losses.append(float(loss_host[(n - 1) % NBUF]))

    def timed(step, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        losses.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(step, n)
        drain_exchange()
        e1.record()
        torch.cuda.synchronize()
        assert PROBE or (len(losses) == n and all(math.isfinite(x) for x in losses))
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    n = max(10, min(args.steps, 400))
    run(step_abi, RING)
    if not args.no_graph:
        drain_exchange()
        torch.cuda.synchronize()
        for k in range(RING):
            replays[k] = sets[k]["step"].capture(frames[k], stage[k % NBUF]["G"], loss_weight=stage[k % NBUF]["G"], stages=False)
        run(step_abi, RING)
    ms = timed(step_abi, n)
    for s in sets:
        s["step"].check_capacity()
    out = {"value": world * n / (ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "steps": n, "ms_per_step": ms / n,
           "pipeline": "inputs prefetched one step ahead on a copy stream; loss read back one step late; "
                       "all copies inside the timed region",
           "api": "C ABI (include/sings_b200.h) driven by sings_b200.step.AvatarStep with preallocated buffers"
                  + ("" if args.no_graph else ", one CUDA graph per frame")}
    if not args.no_graph and not PROBE:
        # Variant: the per-step image goes up as uint8 (3.1 MB instead of 12.6 MB) and is decoded
        # on the device inside the graph.  Reported beside the float32 headline because the
        # float upload is bound by the box's host link, not by the GPU (tools/e2e_probe.sh).
        upload["u8"] = True
        drain_exchange()
        torch.cuda.synchronize()

        def decode(k):
            sb = stage[k % NBUF]
            return lambda: torch.mul(torch.sub(sb["T8"].float(), 127.5), 1.0 / U8_SCALE, out=sb["G"])
        for k in range(RING):
            replays[k] = sets[k]["step"].capture(frames[k], stage[k % NBUF]["G"], loss_weight=stage[k % NBUF]["G"],
                                                 prologue=decode(k), stages=False)
        run(step_abi, RING)
        ms8 = timed(step_abi, n)
        upload["u8"] = False
        out["u8_upload"] = {"value": world * n / (ms8 / 1e3), "unit": "frames/s", "steps": n, "ms_per_step": ms8 / n,
                            "h2d_bytes_per_step": h2d - 3 * H_IMG * W_IMG * 3, "d2h_bytes_per_step": 4,
                            "note": "same C-ABI loop; dL/dimage uploaded as uint8 and decoded on the device "
                                    "(torch sub/mul inside the graph)"}
    if not args.no_dropin:
        nd = max(10, min(args.steps, 100))
        run(step_dropin, RING)   # checked mode: sizes the pair-list capacity for every avatar of the ring
        R.set_async(True)        # then no mid-step host sync: the overflow flag is examined at the next forward
        run(step_dropin, RING)
        msd = timed(step_dropin, nd)
        R.check_pending(block=True)
        R.set_async(False)
        out["dropin"] = {"value": world * nd / (msd / 1e3), "unit": "frames/s", "steps": nd, "ms_per_step": msd / nd,
                         "api": "sings_b200.deform.pose_to_A + deform_gaussians + "
                                "diff_gaussian_rasterization.GaussianRasterizer (torch.autograd)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-dropin", action="store_true", help="skip the autograd drop-in variant of the e2e run")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()

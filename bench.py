#!/usr/bin/env python
"""bench.py -- deformed-avatar frames/s (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W               # this repo's CUDA path, config c2
    torchrun ... bench.py --gpus N --steps K --warmup W         # one rank per GPU, views sharded
    python bench.py --impl reference --steps K --warmup W       # CPU arm (oracle port)
    python bench.py --config c1|c2|c3|c4|c5|shipped             # the other north-star workloads

A step = one pass of the hot path over one synthetic frame: pose -> A -> LBS of every Gaussian
-> rasterize -> dL/dimage -> rasterizer backward -> LBS backward -> densification statistics
(+ the gradient all-reduce when N > 1); configs c1 and c4 (animation) are forward only.
Default workload = BASELINE.json configs[1] (c2): 200k Gaussians, SH degree 3, 1024^2, J = 24.

`value`  : device-resident inputs, AvatarStep (sync-free launch sequence), CUDA events.  For N > 1
           it is the SYNCHRONOUS data-parallel step (the all-reduce of a step's gradients
           finishes before the next forward starts, as an optimizer step needs); the pipelined
           figure (exchange overlapping the next frames) is reported beside it as `dp.pipelined`.
`e2e`    : the same through the C ABI / the public autograd API with HOST buffers: per step
           pose/transl/dL-dimage are copied from pinned memory and the loss is read back, all
           inside the timed region; the box's measured pinned H2D rate is reported with it.
`ab`     : stock comparators on the same box: cub::DeviceRadixSort::SortPairs vs
           sgs_sort_pairs_u64 on the frame's keys; the reference's LBS as eager torch ops on the
           GPU vs sgs_lbs_fwd + sgs_lbs_bwd.
L2       : inputs rotate over a ring of distinct avatars larger than the 126 MB L2.
Nothing here reads /root/reference.  The CPU arm / cpu_baseline / ab.lbs time the oracle port
(oracle/lbs_oracle.py + oracle/c/raster_oracle.c, OpenMP) -- the only use of oracle/ here.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

RING = 4
U8_SCALE = 42.0       # uint8 target = rint(G * 42 + 127.5): +-3 sigma of the N(0,1) gradient image in 8 bits
METRIC = "deformed-avatar fwd+bwd frames/s @200k Gaussians 1024^2"

# BASELINE.json configs (c1..c5) + the configuration the reference ships (human_complex.yaml:34,54,86)
CONFIGS = {
    "c1": dict(workload="CPU torch reference: SMPL LBS of 50k synthetic Gaussians + forward splat of one 512x512 view "
                        "(neutral pose, SH degree 0)",
               N=50_000, H=512, W=512, D=0, J=24, iso=False, neutral=True, mode="fwd"),
    "c2": dict(workload="single B200: LBS + rasterize fwd+bwd, 200k Gaussians, SH degree 3, 1024x1024 view, random pose",
               N=200_000, H=1024, W=1024, D=3, J=24, iso=False, mode="train"),
    "c3": dict(workload="human_complex-style training step: 300k Gaussians, batch of 8 synthetic turn-around views "
                        "sharded over 8xB200 with NCCL grad allreduce",
               N=300_000, H=1024, W=1024, D=3, J=24, iso=False, mode="train", views=8),
    "c4": dict(workload="animation render: 120-frame synthetic AMASS-shaped pose sequence, 200k Gaussians at 1080p, "
                        "frames sharded across 1/2/4/8 GPUs (forward only)",
               N=200_000, H=1080, W=1920, D=3, J=24, iso=False, mode="fwd", frames=120),
    "c5": dict(workload="stress: 1M Gaussians after densification, 2048x2048 views, fwd+bwd, sort and "
                        "atomic-contention scaling",
               N=1_000_000, H=2048, W=2048, D=3, J=24, iso=False, mode="train"),
    "shipped": dict(workload="the configuration SinGS ships (human_complex.yaml): SMPL-H 52 joints, SH degree 0 "
                             "(16 stored coefficients), isotropic Gaussians, 512x896 view, 200k Gaussians, fwd+bwd",
                    N=200_000, H=896, W=512, D=0, J=52, iso=True, mode="train"),
}


def config_dict(cfg):
    """The `config` entry of the bench line: identical in both arms (the driver compares them)."""
    return {"workload": cfg["workload"], "gaussians": cfg["N"], "image": [cfg["H"], cfg["W"]],
            "sh_degree": cfg["D"], "joints": cfg["J"], "isotropic": bool(cfg["iso"]),
            "pass": "forward only" if cfg["mode"] == "fwd" else "forward + backward"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_frame_inputs(cfg, seed: int, yaw: float = 0.0):
    """numpy inputs of one avatar + one frame."""
    from sings_b200 import synthetic as syn
    n, H, W, J = cfg["N"], cfg["H"], cfg["W"], cfg["J"]
    av = syn.make_avatar(n, J, seed=seed, isotropic=cfg["iso"])
    pose = syn.random_pose(J, seed=seed + 2, neutral=cfg.get("neutral", False))
    transl = syn.default_transl(H)
    view = syn.make_view(H, W, yaw=yaw, centre=(0.0, 0.0, float(transl[2])))
    rng = np.random.default_rng(seed + 3)
    G = rng.normal(size=(3, H, W)).astype(np.float32)
    bg = np.ones(3, np.float32)
    return av, pose, transl, view, G, bg


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference torch LBS restated + C rasterizer), all host threads
# ------------------------------------------------------------------------------------------
def cpu_frame(cfg, av, pose, transl, view, G, bg):
    import torch
    from oracle import lbs_oracle as lo
    from oracle import raster_oracle as ro
    t = torch.from_numpy
    train = cfg["mode"] == "train"
    xyz_c = t(av.xyz_canon).requires_grad_(train)
    sc_c = t(av.scales).requires_grad_(train)
    rot_c = None if cfg["iso"] else t(av.rotmat_canon).requires_grad_(train)
    pose_t = t(pose)[None].clone().requires_grad_(train)
    tr = t(transl)[None].clone().requires_grad_(train)
    A = lo.pose_to_A(pose_t, t(av.rest), av.parents, t(av.inv_A_t2cano))
    xyz, q, sc, _ = lo.deform(A, xyz_c, t(av.lbs_weights), sc_c, rot_c, None, tr)
    cam = ro.Camera(W=view.image_width, H=view.image_height, tanfovx=view.tanfovx, tanfovy=view.tanfovy,
                    view=view.world_view_transform.reshape(-1), proj=view.full_proj_transform.reshape(-1),
                    campos=view.camera_center)
    st = ro.forward(cam, xyz[0].detach().numpy(), av.opacity, bg, shs=av.shs,
                    scales=sc[0].detach().numpy(), rotations=q[0].detach().numpy(), sh_degree=cfg["D"])
    if train:
        gr = ro.backward(st, G)
        torch.autograd.backward([xyz, q, sc], [t(gr["means3D"])[None], t(gr["rotations"])[None], t(gr["scales"])[None]])
    return st.num_rendered, float((st.color * G).sum())


def run_cpu(cfg, steps: int, warmup: int, frames):
    import torch
    from oracle import raster_oracle as ro
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ro.set_threads(cores)
    for i in range(warmup):
        cpu_frame(cfg, *frames[i % len(frames)])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_frame(cfg, *frames[i % len(frames)])
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, cores


def cpu_sample_note(cfg, n):
    what = "fwd+bwd" if cfg["mode"] == "train" else "forward only"
    return (f"{n} full frames of the same workload on the host ({what}): torch restatement of the reference's "
            f"lbs_extra / rotations + OpenMP C rasterizer oracle (the reference has no CPU rasterizer and its CUDA "
            f"one is not vendored)")


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = [build_frame_inputs(cfg, 100)]
    # bound the run to a few minutes: probe one frame, then cap the step count
    t0 = time.perf_counter()
    cpu_frame(cfg, *frames[0])
    t1 = time.perf_counter() - t0
    steps_run = max(1, min(args.steps, int(150.0 / max(t1, 1e-3))))
    warm_run = min(args.warmup, 1)
    fps, ms, cores = run_cpu(cfg, steps_run, warm_run, frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps_run, "warmup": warm_run, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_note(cfg, steps_run)},
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
def algorithmic_bytes(N, J, M, L, W, H, n_vis, iso, passes):
    """Per-launch algorithmic bytes of each stage (DESIGN.md 'Algorithmic bytes')."""
    pix, tiles = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    rot = 0 if iso else 36
    return {
        "lbs_fwd": N * (64 + rot + 4 * J),
        "lbs_bwd": N * (112 + rot + 4 * J),
        "geometry": N * (236 + 75),                                 # preprocessCUDA
        # InclusiveSum + duplicateWithKeys + SortPairs + identifyTileRanges: one stage
        "binning": 8 * N + 20 * N + 12 * L + (8 + 24 * passes) * L + 8 * L + 8 * tiles,
        "blend_fwd": 40 * L + 20 * pix + 8 * tiles,
        "blend_bwd": 40 * L + 20 * pix + 36 * n_vis,
        "geometry_bwd": (300 + 256) * n_vis,
    }


def measure_h2d(dev, nbytes):
    """Pinned host -> device copy rate of THIS box for a buffer the size of the per-step upload."""
    import torch
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    return nbytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9


def ab_baselines(cfg, s, dev):
    """Stock comparators on this box (rank 0, 1 GPU): CUB SortPairs, eager-torch LBS."""
    import torch
    from sings_b200 import _lib, deform
    out = {}
    L_ = _lib.lib()
    st = torch.cuda.current_stream(dev).cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            a, b = ev(), ev()
            a.record(); fn(); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    # ---- sort: the frame's keys (tile|depth, 45-47 bits), shuffled, with their ids ----
    try:
        step = s["step"]
        from sings_b200.rasterizer import layout_info
        info = layout_info(step.N, step.Wd, step.H, step.L_cap)
        L = int(s["L"])
        raw = step.binning
        keys_sorted = raw[info["keys_sorted"]:info["keys_sorted"] + 8 * L].view(torch.int64).clone()
        perm = torch.randperm(L, device=dev, generator=torch.Generator(dev).manual_seed(1))
        k0 = keys_sorted[perm].contiguous()
        v0 = perm.to(torch.int32).contiguous()
        end_bit = info["end_bit"]
        kin, vin = k0.clone(), v0.clone()
        kt, vt = torch.empty_like(k0), torch.empty_like(v0)
        sb = int(L_.sgs_sort_scratch_bytes(L))
        scratch = torch.empty(sb, device=dev, dtype=torch.uint8)
        flag = C.c_int(0)

        def ours():
            kin.copy_(k0); vin.copy_(v0)
            _lib.check(L_.sgs_sort_pairs_u64(kin.data_ptr(), vin.data_ptr(), kt.data_ptr(), vt.data_ptr(),
                                             scratch.data_ptr(), sb, L, end_bit, C.byref(flag), st), "sort")

        def copy_only():
            kin.copy_(k0); vin.copy_(v0)
        t_copy = timed(copy_only)
        t_ours = timed(ours) - t_copy
        cub_path = os.path.join(ROOT, "tools", "_build", "libcub_ab.so")
        if os.path.exists(cub_path):
            cub = C.CDLL(cub_path)
            cub.cub_ab_sort_pairs_u64.restype = C.c_int
            cub.cub_ab_sort_pairs_u64.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_int, C.c_void_p,
                                                                    C.POINTER(C.c_size_t), C.c_void_p]
            tb = C.c_size_t(0)
            cub.cub_ab_sort_pairs_u64(None, None, None, None, L, end_bit, None, C.byref(tb), st)
            temp = torch.empty(max(int(tb.value), 1), device=dev, dtype=torch.uint8)

            def theirs():
                rc = cub.cub_ab_sort_pairs_u64(k0.data_ptr(), kt.data_ptr(), v0.data_ptr(), vt.data_ptr(), L, end_bit,
                                               temp.data_ptr(), C.byref(tb), st)
                assert rc == 0
            t_cub = timed(theirs)
            theirs()
            ck, cv = kt.clone(), vt.clone()
            ours()
            torch.cuda.synchronize(dev)
            rk, rv = (kt, vt) if flag.value else (kin, vin)
            same = bool(torch.equal(rk, ck) and torch.equal(rv, cv))
            out["sort"] = {"pairs": L, "end_bit": end_bit, "cub_ms": round(t_cub, 4), "ours_ms": round(t_ours, 4),
                           "speedup_vs_cub": round(t_cub / t_ours, 3), "identical_output": same,
                           "note": "cub::DeviceRadixSort::SortPairs(u64, u32) vs sgs_sort_pairs_u64 on the frame's "
                                   "shuffled keys, whole sort incl. histogram; in the frame itself the four depth digits "
                                   "are sorted per Gaussian, not per pair, so only two of these passes run over the pairs"}
        else:
            out["sort"] = {"unavailable": "tools/_build/libcub_ab.so not built", "ours_ms": round(t_ours, 4)}
    except Exception as e:       # a failed comparator must not take the bench line down
        out["sort"] = {"unavailable": repr(e)[:200]}

    # ---- LBS: the reference's torch ops (restated in oracle/lbs_oracle.py) run eagerly on the GPU ----
    try:
        from oracle import lbs_oracle as lo
        av = s["av"]
        t = lambda a: torch.as_tensor(a, device=dev)
        A = lo.pose_to_A(t(s["pose"])[None], t(av.rest), av.parents, t(av.inv_A_t2cano))
        Wd = t(av.lbs_weights)
        xyz = t(av.xyz_canon).requires_grad_(True)
        scl = t(av.scales).requires_grad_(True)
        rot = None if cfg["iso"] else t(av.rotmat_canon).requires_grad_(True)
        tr = t(s["transl"])[None]
        n = cfg["N"]
        gx, gq, gs = torch.randn(1, n, 3, device=dev), torch.randn(1, n, 4, device=dev), torch.randn(1, n, 3, device=dev)

        def torch_lbs():
            for p in (xyz, scl, rot):
                if p is not None:
                    p.grad = None
            x, q, sc, _ = lo.deform(A, xyz, Wd, scl, rot, None, tr)
            torch.autograd.backward([x, q, sc], [gx, gq, gs])

        def our_lbs():
            for p in (xyz, scl, rot):
                if p is not None:
                    p.grad = None
            x, q, sc = deform.deform_gaussians(A[0], xyz, Wd, rot, scl, None, tr[0])
            torch.autograd.backward([x, q, sc], [gx[0], gq[0], gs[0]])
        t_torch, t_ours = timed(torch_lbs, 10), timed(our_lbs, 10)
        out["lbs"] = {"torch_eager_ms": round(t_torch, 4), "ours_ms": round(t_ours, 4),
                      "speedup_vs_torch": round(t_torch / t_ours, 2), "kind": "port",
                      "note": "deform segment fwd+bwd (lbs_extra + compose + matrix_to_quaternion, autograd) as eager "
                              "torch CUDA ops vs sings_b200.deform.deform_gaussians (sgs_lbs_fwd + sgs_lbs_bwd through "
                              "autograd); the fused per-frame path folds both into the rasterizer's kernels"}
    except Exception as e:
        out["lbs"] = {"unavailable": repr(e)[:200]}

    # ---- image loss: the reference's l1_loss + ssim (restated in oracle/loss_oracle.py) run eagerly on the GPU ----
    try:
        from oracle import loss_oracle as llo
        from sings_b200.losses import image_loss
        H, W = cfg["H"], cfg["W"]
        gen = torch.Generator(dev).manual_seed(3)
        gt = torch.rand(3, H, W, device=dev, generator=gen)
        pred = (gt + 0.1 * torch.randn(3, H, W, device=dev, generator=gen)).clamp(0, 1).requires_grad_(True)
        mask = (torch.rand(H, W, device=dev, generator=gen) > 0.3).float()
        bgc = torch.ones(3, device=dev)

        def torch_loss():
            pred.grad = None
            llo.human_image_loss(pred, gt, mask, bgc)[0].backward()

        def our_loss():
            pred.grad = None
            image_loss(pred, gt, mask, bgc)[0].backward()
        t_torch, t_ours = timed(torch_loss, 10), timed(our_loss, 10)
        out["loss"] = {"torch_eager_ms": round(t_torch, 4), "ours_ms": round(t_ours, 4),
                       "speedup_vs_torch": round(t_torch / t_ours, 2), "kind": "port",
                       "note": "L1 + SSIM image loss fwd+bwd (HumanLoss.forward's image terms: l1_loss + ssim, five "
                               "depthwise conv2d + elementwise ops, autograd) as eager torch CUDA ops vs "
                               "sings_b200.losses.image_loss (sgs_image_loss_fwd + _bwd through autograd)"}
    except Exception as e:
        out["loss"] = {"unavailable": repr(e)[:200]}

    # ---- neighbour search of the scale-edge loss (loss_items.py:57-90; pytorch3d.knn_points is not installed:
    #      the comparator is the same exact search as chunked torch.cdist + topk on the GPU) ----
    try:
        from sings_b200.losses import knn_points
        xyz = torch.as_tensor(s["av"].xyz_canon, device=dev).float().contiguous()

        def torch_knn(rows=None):
            outs = []
            n_rows = xyz.shape[0] if rows is None else rows
            for i in range(0, n_rows, 4096):
                d = torch.cdist(xyz[i:min(i + 4096, n_rows)], xyz)       # (matmul formulation: fast, ~1e-3 relative)
                outs.append(torch.topk(d, 9, dim=1, largest=False)[0][:, 1:].mean(1))
            return torch.cat(outs)
        t_ours = timed(lambda: knn_points(xyz, 8), 10)
        for _ in range(2):
            torch_knn()
        torch.cuda.synchronize(dev)
        a, b = ev(), ev()
        a.record(); ref = torch_knn(); b.record()
        torch.cuda.synchronize(dev)
        t_torch = a.elapsed_time(b)
        # exactness on a sample of rows with the direct (non-matmul) distance
        d = torch.cdist(xyz[:2048], xyz, compute_mode="donot_use_mm_for_euclid_dist")
        exact = torch.topk(d, 9, dim=1, largest=False)[0][:, 1:].mean(1)
        same = bool(torch.allclose(knn_points(xyz, 8)[:2048], exact, rtol=1e-4, atol=1e-7))
        out["knn"] = {"points": int(xyz.shape[0]), "K": 8, "torch_cdist_topk_ms": round(t_torch, 3), "ours_ms": round(t_ours, 4),
                      "speedup_vs_torch": round(t_torch / t_ours, 1), "same_mean_distances": same, "kind": "port",
                      "note": "mean distance to the 8 nearest other Gaussians (GaussiansEdgeLoss, every iteration in the "
                              "reference): exact grid search (sgs_knn_mean_dist) vs brute-force chunked torch.cdist + topk"}
    except Exception as e:
        out["knn"] = {"unavailable": repr(e)[:200]}

    # ---- tri-plane features (hexplane.py, shipped configuration: 32 channels, 64^3, multires 1/2/4) ----
    try:
        from oracle import hexplane_oracle as hpo
        from sings_b200.triplane import HexPlaneField
        hcfg = {"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": 32, "resolution": [64, 64, 64],
                "multires": [1, 2, 4]}
        field = HexPlaneField(hcfg, bounds=1.3, device=dev)
        params = [p for gp in field.grids for p in gp]
        pts = torch.as_tensor(s["av"].xyz_canon, device=dev).float().contiguous().requires_grad_(True)
        d_out = torch.randn(pts.shape[0], 96, device=dev)

        def clear():
            pts.grad = None
            for p in params:
                p.grad = None

        def torch_hex():
            clear()
            f = hpo.hexplane_features(pts, field.aabb, [params[3 * k:3 * k + 3] for k in range(3)])
            (f * d_out).sum().backward()

        def our_hex():
            clear()
            (field(pts) * d_out).sum().backward()
        t_torch, t_ours = timed(torch_hex, 10), timed(our_hex, 10)
        out["hexplane"] = {"points": int(pts.shape[0]), "torch_eager_ms": round(t_torch, 4), "ours_ms": round(t_ours, 4),
                           "speedup_vs_torch": round(t_torch / t_ours, 2), "kind": "port",
                           "note": "HexPlaneField.forward + backward over all Gaussians (nine grid_sample launches, products, "
                                   "concatenation, autograd) as eager torch CUDA ops vs sings_b200.triplane.HexPlaneField "
                                   "(sgs_hexplane_fwd + _bwd on channel-last planes, incl. the layout copies)"}
    except Exception as e:
        out["hexplane"] = {"unavailable": repr(e)[:200]}
    # ---- non-image loss terms of gs_trainer.py:363-396: region Laplacians (position, colour, hands) + L2Norm ----
    try:
        from oracle import reg_oracle as rgo
        from sings_b200.regularizers import L2Norm, RegionLaplacianLoss_v2
        gen = torch.Generator().manual_seed(7)
        V, R_ = 110_000, 15                    # the subdivided SMPL template the reference trains on
        labels = torch.sort(torch.randint(0, R_, (V,), generator=gen)).values
        a = torch.arange(V).repeat_interleave(3)
        b_ = (a + torch.randint(1, 40, (3 * V,), generator=gen)).clamp(max=V - 1)
        edges = torch.unique(torch.sort(torch.stack([a, b_], 1), dim=1).values, dim=0)
        edges = edges[edges[:, 0] != edges[:, 1]].to(dev)
        labels = labels.to(dev)
        w = np.linspace(0.5, 1.5, R_)
        verts = torch.randn(V, 3, device=dev)
        ours_lap = RegionLaplacianLoss_v2(verts, edges, labels, region_weights=w)
        ref_lap = rgo.RegionLaplacian(verts, edges, labels, w, sparse=True)
        N = int(s["av"].xyz_canon.shape[0])
        xa = torch.randn(V, 3, device=dev, requires_grad=True)
        sh = torch.randn(V, 16, 3, device=dev, requires_grad=True)
        off = (0.01 * torch.randn(N, 3, device=dev)).requires_grad_(True)
        sc = (0.002 + 0.01 * torch.rand(N, 1, device=dev)).repeat(1, 3).requires_grad_(True)
        opa = torch.rand(N, 1, device=dev, requires_grad=True)
        l2cfg = dict(lambda_xyz_offsets=0.001, lambda_scales_diff=0.005, max_scale_threshold=0.005, lambda_max_scale=0.01,
                     min_opacity_threshold=0.2, lambda_min_opacity=0.001)
        ours_l2 = L2Norm(**l2cfg)
        leaves = (xa, sh, off, sc, opa)

        def run(lap, l2):
            for t in leaves:
                t.grad = None
            d = {"xyz_offsets": off, "scales": sc, "opacity": opa}
            loss = 500.0 * lap.forward(xa) + 5.0 * lap.forward(sh[:, 0]) + 1e-5 * lap.forward_hands(xa) + l2(d)
            loss.backward()
            return loss
        l_ref = float(run(ref_lap, lambda d: rgo.l2norm(d, **l2cfg)))
        l_our = float(run(ours_lap, ours_l2))
        t_torch = timed(lambda: run(ref_lap, lambda d: rgo.l2norm(d, **l2cfg)), 10)
        t_ours = timed(lambda: run(ours_lap, ours_l2), 10)
        out["regularizers"] = {"vertices": V, "gaussians": N, "torch_eager_ms": round(t_torch, 4), "ours_ms": round(t_ours, 4),
                               "speedup_vs_torch": round(t_torch / t_ours, 2), "kind": "port",
                               "same_loss": bool(abs(l_ref - l_our) <= 1e-4 * abs(l_ref)),
                               "note": "RegionLaplacianLoss_v2 on positions, colours and hands + L2Norm, fwd+bwd (per region a "
                                       "torch.sparse matmul, pow, mean; boolean-mask norms) as eager torch CUDA ops vs "
                                       "sings_b200.regularizers (one CSR operator; sgs_laplacian_loss_* and sgs_l2norm_*)"}
    except Exception as e:
        out["regularizers"] = {"unavailable": repr(e)[:200]}
    return out


def gpu_arm(args, cfg):
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
    N, H, W, D, J = cfg["N"], cfg["H"], cfg["W"], cfg["D"], cfg["J"]
    train = cfg["mode"] == "train"
    ring = RING if N <= 400_000 else 2          # 1M-Gaussian avatars: two are already > L2 several times over
    views = cfg.get("views", 0)

    # ---- workload: ring of distinct avatars (inputs larger than L2), each rank its own views / frames
    sets = []
    for r in range(ring):
        yaw = 2 * math.pi * ((rank * ring + r) % views) / views if views else 0.0
        av, pose, transl, view, G, bg = build_frame_inputs(cfg, 1000 * rank + 10 * r, yaw=yaw)
        t = lambda a: torch.as_tensor(a, device=dev)
        step = AvatarStep(t(av.xyz_canon), None if cfg["iso"] else t(av.rotmat_canon), t(av.scales), t(av.opacity),
                          t(av.shs), t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano),
                          H, W, D, timing=True)
        step.forward_only = not train          # c1 / c4: no backward follows, the forward leaves nothing for one
        fr = FrameInputs(pose=t(pose), transl=t(transl), viewmatrix=t(view.world_view_transform),
                         projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(bg),
                         tanfovx=view.tanfovx, tanfovy=view.tanfovy)
        sets.append(dict(step=step, fr=fr, G=t(G), av=av, pose=pose, transl=transl, view=view, G_np=G, bg=bg))
    exch = dp.GradExchange(N, sets[0]["step"].n_param_grads, dev,
                           defer_max=not os.environ.get("SGS_DP_MAX_EVERY_STEP")) if (world > 1 and train) else None

    def one_step(i, pending, staged=False, sync_dp=False):
        s = sets[i % ring]
        st = s["step"]
        st.record_stages = bool(staged)    # eager launches: stage events only in the instrumented passes
        if exch is not None and pending[i % ring] is not None:
            pending[i % ring]()            # finish the all-reduce that last used this bucket (folds + clears the step statistics)
            pending[i % ring] = None
        if "replay" in s:
            s[{False: "replay", True: "replay_staged", "coarse": "replay_coarse"}[staged]]()   # the frame as one CUDA-graph launch
        else:
            st.forward(s["fr"])
            if train:
                st.backward(s["G"])
        if exch is not None:
            fin = exch.exchange(st.bucket, st.max_radii2D, async_op=True, reset_step=True)
            if sync_dp:
                fin()                      # the current stream waits for the reduced bucket: nothing of the next step overlaps it
            else:
                pending[i % ring] = fin

    def drain(pending):
        for k in range(ring):
            if pending[k] is not None:
                pending[k]()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(n, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            one_step(i, pending, **kw)
        drain(pending)
        if exch is not None:
            exch.sync_max()                # the deferred all-reduce(MAX) of max_radii2D, inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    pending = [None] * ring
    from sings_b200._lib import SgsError
    for attempt in range(4):
        for i in range(max(args.warmup, 3, ring)):
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
        # capacities are settled: record each avatar's frame (forward [+ backward]) as a CUDA graph
        drain(pending)
        for s in sets:
            cap = (lambda s, **kw: s["step"].capture(s["fr"], s["G"] if train else None, **kw))
            s["replay"] = cap(s, stages=False)
            s["replay_staged"] = cap(s, stages=True)
            # coarse: only the events around LBS + preprocess + sort + ranges (8 = frame start, 3 = after ranges)
            s["replay_coarse"] = cap(s, stages=True, stage_mask=(1 << 8) | (1 << 3))
        for i in range(2 * ring):
            one_step(i, pending)
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dp_info = None
    if exch is not None:
        ms_pipe = timed_region(args.steps)                       # exchange overlapping the next frames
        ms_total = timed_region(args.steps, sync_dp=True)        # synchronous step: the headline for N > 1
        dp_info = {"mode": "synchronous (all-reduce of a step finishes before the next forward starts)",
                   "pipelined": {"value": world * args.steps / (ms_pipe / 1e3), "ms_per_step": ms_pipe / args.steps,
                                 "note": f"the all-reduce of a step overlaps the next {ring - 1} frames (distinct avatars)"},
                   "bucket_mb": round(sets[0]["step"].bucket.numel() * 4 / 1e6, 1),
                   "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")}
    else:
        ms_total = timed_region(args.steps)
    for s in sets:
        s["step"].check_capacity()
    value = world * args.steps / (ms_total / 1e3)

    # ---- per-stage device times: a second timed pass over the same frames with CUDA events
    # recorded at the stage boundaries.  The records sit between kernels (graph nodes), which
    # turns the kernels' overlapped programmatic launch into full dependencies -- ~4 us per
    # event -- so they stay out of the headline region above; the stage times below therefore sum
    # to more than ms_per_step.
    n_stage = max(ring, min(args.steps, 100))
    ms_staged = timed_region(n_stage, staged=True) / n_stage
    stage = {}
    for s in sets:
        for k, v in s["step"].stage_ms(train).items():
            stage.setdefault(k, []).append(v)
    # third pass: one interval over the north star's target set (LBS + preprocess + sort + ranges),
    # two event records instead of five inside it -- less perturbation than the sum of its stages
    hot_coarse = None
    if not args.no_graph:
        timed_region(n_stage, staged="coarse")
        hot_coarse = float(np.mean([s["step"].interval_ms(8, 3) for s in sets]))
    stage = {k: float(np.mean(v)) for k, v in stage.items()}
    Lm = float(np.mean([s["L"] for s in sets]))
    n_vis = float(np.mean([int((s["step"].radii > 0).sum().item()) for s in sets]))
    from sings_b200.rasterizer import layout_info
    passes = layout_info(N, W, H, 1)["passes"]
    ab = algorithmic_bytes(N, J, 16, Lm, W, H, n_vis, cfg["iso"], passes)
    if "deform_geometry" in stage:      # fused kernels: the stages merge, their algorithmic bytes add
        ab["deform_geometry"] = ab.pop("lbs_fwd") + ab.pop("geometry")
        ab["geometry_lbs_bwd"] = ab.pop("geometry_bwd") + ab.pop("lbs_bwd")
    if not train:
        for k in ("blend_bwd", "geometry_bwd", "lbs_bwd", "geometry_lbs_bwd"):
            ab.pop(k, None)
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
                "binning": "emit_pairs_kernel", "blend_fwd": "blend_fwd_kernel",
                "blend_bwd": "blend_bwd_kernel", "geometry_bwd": "geometry_bwd_kernel"}[dom]
        if kern in tj.get("kernels", {}) and tj.get("config", "c2") == args.config:
            traffic, traffic_src = tj["kernels"][kern]["dram_bytes"], f"profiles/{tj.get('source')} ({kern})"
            if dom == "binning":        # a stage of several kernels: DRAM bytes of all its launches, not of the emission alone
                try:
                    parts = [("onesweep_pass_kernel", 4 + (passes - 4)), ("emit_pairs_kernel", 1), ("tile_ranges_kernel", 1)]
                    traffic = int(sum(tj["kernels"][k]["dram_bytes"] * n for k, n in parts))
                    traffic_src = f"profiles/{tj.get('source')} (" + " + ".join(f"{n} x {k}" for k, n in parts) + ")"
                except Exception:
                    pass
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
                                          if hot_coarse else "sum of the stage intervals"},
        "frame_alg_mb": round(sum(ab.values()) / 1e6, 1),
        "pairs_L": Lm, "visible": n_vis,
        "stage_timing": {"steps": n_stage, "ms_per_step": round(ms_staged, 4),
                         "note": "stage times come from a second timed pass of the same frames with CUDA events at the "
                                 "stage boundaries; the event records break the kernels' programmatic dependent launch "
                                 "(~4 us each), so that pass is slower than the headline region, which has none"},
    }

    # ---- informational: two independent frames in flight (two streams, each replaying the graphs of its own
    # avatars).  NOT the headline: a training step with one view per step is serial by the optimizer.  It applies
    # where frames really are independent -- animation (c4), several views per rank and step (c3 below 8 GPUs) --
    # and shows how much of the frame is latency rather than throughput: the per-Gaussian and binning kernels of
    # one frame leave the SMs partly idle, a second frame's kernels fill them.
    concurrent = None
    if rank == 0 and world == 1 and not args.no_graph and ring >= 2 and ring % 2 == 0:
        try:
            lanes = [torch.cuda.Stream(dev) for _ in range(2)]
            n_c = max(2 * ring, min(args.steps, 200))

            def concurrent_region(n):
                torch.cuda.synchronize(dev)
                cur = torch.cuda.current_stream(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
                for ln in lanes:
                    ln.wait_event(e0)
                for i in range(n):                      # avatar i % ring always runs on lane i % 2: its buffers never cross lanes
                    with torch.cuda.stream(lanes[i % 2]):
                        sets[i % ring]["replay"]()
                for ln in lanes:
                    done = torch.cuda.Event()
                    done.record(ln)
                    cur.wait_event(done)
                e1.record(cur)
                torch.cuda.synchronize(dev)
                return e0.elapsed_time(e1)
            concurrent_region(2 * ring)
            ms_c = concurrent_region(n_c)
            for s_ in sets:
                s_["step"].check_capacity()
            concurrent = {"frames_in_flight": 2, "value": n_c / (ms_c / 1e3), "unit": "frames/s", "steps": n_c,
                          "ms_per_frame": ms_c / n_c, "vs_serial": round((n_c / (ms_c / 1e3)) / value, 3),
                          "note": "two streams, each replaying the frame graphs of its own avatars; informational (independent "
                                  "frames only: animation, several views per rank and step), not the headline"}
        except Exception as e:       # a failed probe must not take the bench line down
            concurrent = {"unavailable": repr(e)[:200]}
        torch.cuda.synchronize(dev)

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, cfg, sets, dev, world, rank, exch, ring)
    clocks = sampler.stop() if rank == 0 else None

    ab_out = None
    if rank == 0 and world == 1 and not args.no_ab:
        ab_out = ab_baselines(cfg, sets[0], dev)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        s = sets[0]
        frames = [(s["av"], s["pose"], s["transl"], s["view"], s["G_np"], s["bg"])]
        # bounded sample of the same workload: ~15 s of host work (probe one frame, then size the run)
        t0 = time.perf_counter()
        cpu_frame(cfg, *frames[0])
        n_cpu = max(2, min(200, int(15.0 / max(time.perf_counter() - t0, 1e-3))))
        fps, ms, cores = run_cpu(cfg, n_cpu, 0, frames)
        cpu_base = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": cpu_sample_note(cfg, n_cpu)}
    if rank == 0:
        st0 = sets[0]["step"]
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg),
            "run": {"config_name": args.config, "views_per_gpu_per_step": 1,
                    "launch": "eager launches" if args.no_graph else "one CUDA graph per frame",
                    "l2": f"inputs rotate over a ring of {ring} distinct avatars (~{ring * (70 if N <= 400_000 else 350)} MB of "
                          f"inputs) > 126 MB L2",
                    "deformer": f"fused into the rasterizer's per-Gaussian kernels, skinning weights packed to {st0.K} slots"
                                if st0.K else "separate LBS kernels (dense skinning weights)",
                    "parallelism": f"dp{world} (views sharded, gradient bucket all-reduced every step)" if exch is not None
                                   else (f"dp{world} (frames sharded, no collective)" if world > 1 else "single GPU")},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks, "dp": dp_info, "ab": ab_out,
            "concurrent": concurrent,
            "gpu_launches": st0.launches_per_frame(train) * args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, cfg, sets, dev, world, rank, exch, ring):
    """Same metric end to end with HOST buffers.  Per step, inside the timed region: H2D of
    pose + transl + dL/dimage from pinned memory, forward, loss, backward, D2H of the loss
    (forward-only configs: H2D of pose + transl, forward, clamp + 8-bit conversion on the
    device, D2H of the finished uint8 frame into pinned memory -- the reference's
    `image.cpu()` + cv2 conversion, gs_trainer.py:716-719).

    Two callers of the same kernels are timed:
      * `e2e` (headline): the C-ABI path -- `AvatarStep` drives include/sings_b200.h directly
        with preallocated buffers (what a trainer integrating the library calls per frame);
      * `e2e.dropin`: the reference-facing autograd modules (`sings_b200.deform` +
        `diff_gaussian_rasterization.GaussianRasterizer`), which add torch.autograd and
        per-call allocation overhead on the host.
    Both use the same input pipeline: inputs are prefetched one step ahead on a copy stream
    into double-buffered device staging (like a data loader with pinned memory) and the result
    is read back through pinned slots one step late (like a logging trainer)."""
    import torch
    import torch.distributed as dist
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from sings_b200 import deform
    from sings_b200 import rasterizer as R
    from sings_b200.step import FrameInputs

    N, H, W, D, J = cfg["N"], cfg["H"], cfg["W"], cfg["D"], cfg["J"]
    train = cfg["mode"] == "train"
    for s in sets:
        s["step"].record_stages = False     # no stage events on this path (eager launches included)
    host = []
    for s in sets:
        t = lambda a: torch.as_tensor(a, device=dev)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        # the same dense gradient image quantised to 8 bits: how a target image is stored on disk
        t8 = np.ascontiguousarray(np.clip(np.rint(s["G_np"] * U8_SCALE + 127.5), 0, 255).astype(np.uint8).transpose(1, 2, 0))
        host.append(dict(pose=pin(s["pose"]), transl=pin(s["transl"]), G=pin(s["G_np"]), T8=pin(t8), view=s["view"],
                         bg=t(s["bg"]), vm=t(s["view"].world_view_transform),
                         pm=t(s["view"].full_proj_transform), cp=t(s["view"].camera_center)))
    h2d = int(host[0]["pose"].numel() * 4 + host[0]["transl"].numel() * 4 + (host[0]["G"].numel() * 4 if train else 0))
    d2h = 4 if train else 3 * H * W
    h2d_gbs = measure_h2d(dev, max(h2d, 1 << 20))
    cur = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    NBUF = 2
    stage = [dict(pose=torch.empty(J, 3, device=dev), transl=torch.empty(3, device=dev),
                  G=torch.empty(3, H, W, device=dev),
                  T8=torch.empty(H, W, 3, device=dev, dtype=torch.uint8), ready=torch.cuda.Event(),
                  free=torch.cuda.Event()) for _ in range(NBUF)]
    loss_host = torch.zeros(NBUF).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(NBUF)]
    writer = {"w": None}        # forward-only configs: the asynchronous frame writer (sings_b200.animate.FrameWriter)
    losses = []
    # diagnostics only (tools/e2e_probe.sh): "nog" skips the dL/dimage upload, "nosync" the lagged loss read
    PROBE = os.environ.get("SGS_E2E_PROBE", "")
    upload = {"u8": False}      # True: the per-step image upload is uint8 and decoded on the device

    def prefetch(i):
        hs, sb = host[i % ring], stage[i % NBUF]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sb["free"])          # the step that last used this buffer is done
            sb["pose"].copy_(hs["pose"], non_blocking=True)
            sb["transl"].copy_(hs["transl"], non_blocking=True)
            if train:
                if upload["u8"]:
                    sb["T8"].copy_(hs["T8"], non_blocking=True)
                elif "nog" not in PROBE:
                    sb["G"].copy_(hs["G"], non_blocking=True)
            sb["ready"].record(copy_stream)

    def finish_step(i, result, sb):
        sb["free"].record(cur)
        if not train:
            # the finished frame: 8-bit conversion on the device, async copy into pinned memory, hand-over
            # to the encoder threads (JPEG with --encode, else the frame is only delivered)
            writer["w"].submit(result, f"{i:05d}")
            return
        loss_host[i % NBUF:i % NBUF + 1].copy_(result.detach().reshape(1), non_blocking=True)
        loss_ev[i % NBUF].record(cur)
        if i >= 1 and "nosync" not in PROBE:           # read the previous step's result
            loss_ev[(i - 1) % NBUF].synchronize()
            losses.append(float(loss_host[(i - 1) % NBUF]))

    # ---- C-ABI path ----
    pending = [None] * ring

    def drain_exchange():
        for k in range(ring):
            if pending[k] is not None:
                pending[k]()
                pending[k] = None

    assert ring % NBUF == 0        # ring slot k always meets staging buffer k % NBUF (graphs bind addresses)
    frames = [FrameInputs(pose=stage[k % NBUF]["pose"], transl=stage[k % NBUF]["transl"], viewmatrix=host[k]["vm"],
                          projmatrix=host[k]["pm"], campos=host[k]["cp"], bg=host[k]["bg"],
                          tanfovx=host[k]["view"].tanfovx, tanfovy=host[k]["view"].tanfovy) for k in range(ring)]
    replays = [None] * ring

    def step_abi(i, n_total):
        sb, st = stage[i % NBUF], sets[i % ring]["step"]
        if i + 1 < n_total:
            prefetch(i + 1)
        cur.wait_event(sb["ready"])
        if replays[i % ring] is not None:
            replays[i % ring]()            # forward + loss + backward as one CUDA-graph launch
            result = st.loss if train else st.color
        else:
            img = st.forward(frames[i % ring])
            if train:
                result = torch.dot(img.view(-1), sb["G"].view(-1))      # L = sum(image * G); dL/dimage = G
                st.backward(sb["G"])
            else:
                result = img
        if exch is not None:
            exch.exchange(st.bucket, st.max_radii2D, async_op=True, reset_step=True)()   # synchronous step (see gpu_arm)
        finish_step(i, result, sb)

    # ---- drop-in autograd path ----
    params = []
    if not args.no_dropin and train:
        for s in sets:
            av = s["av"]
            t = lambda a: torch.as_tensor(a, device=dev)
            params.append(dict(xyz=t(av.xyz_canon).requires_grad_(True),
                               rot=None if cfg["iso"] else t(av.rotmat_canon).requires_grad_(True),
                               scales=t(av.scales).requires_grad_(True), opacity=t(av.opacity).requires_grad_(True),
                               shs=t(av.shs).requires_grad_(True), W=t(av.lbs_weights), rest=t(av.rest),
                               parents=torch.from_numpy(av.parents).to(device=dev, dtype=torch.int32),
                               inv_A=t(av.inv_A_t2cano)))

    def step_dropin(i, n_total):
        hs, sb, p = host[i % ring], stage[i % NBUF], params[i % ring]
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
            bg=hs["bg"], scale_modifier=1.0, viewmatrix=hs["vm"], projmatrix=hs["pm"], sh_degree=D,
            campos=hs["cp"], prefiltered=False, debug=False)
        means2D = torch.zeros_like(xyz, requires_grad=True)
        img, radii = GaussianRasterizer(rs)(means3D=xyz, means2D=means2D, shs=p["shs"], opacities=p["opacity"],
                                            scales=sc, rotations=rotq)
        loss = (img * sb["G"]).sum()
        for q in (p["xyz"], p["rot"], p["scales"], p["opacity"], p["shs"]):
            if q is not None:
                q.grad = None
        loss.backward()
        finish_step(i, loss, sb)

    # ---- one-call autograd path (sings_b200.fused.AvatarRenderer) ----
    renderers = []

    def step_fused(i, n_total):
        hs, sb, p, r = host[i % ring], stage[i % NBUF], params[i % ring], renderers[i % ring]
        if i + 1 < n_total:
            prefetch(i + 1)
        cur.wait_event(sb["ready"])
        pose = sb["pose"].detach().requires_grad_(True)
        transl = sb["transl"].detach().requires_grad_(True)
        v = hs["view"]
        img, radii = r(pose, transl, hs["vm"], hs["pm"], hs["cp"], hs["bg"], v.tanfovx, v.tanfovy)
        loss = (img * sb["G"]).sum()
        for q in (p["xyz"], p["rot"], p["scales"], p["opacity"], p["shs"]):
            if q is not None:
                q.grad = None
        loss.backward()
        finish_step(i, loss, sb)

    def run(step, n):
        for sb in stage:
            sb["free"].record(cur)
        if not train:
            from sings_b200.animate import FrameWriter
            out_dir = os.path.join(ROOT, "gpurun_out", f"anim_rank{rank}") if args.encode else None
            writer["w"] = FrameWriter(out_dir, H, W, dev, depth=8, workers=8, ext="jpg", encode=args.encode,
                                      sink=lambda name, arr: losses.append(float(arr[0, 0, 0])))
        prefetch(0)
        for i in range(n):
            step(i, n)
        if not train:
            assert writer["w"].close() == n                # every frame copied out (and encoded) inside the timed region
            if args.encode:
                losses.extend([0.0] * n)
            return
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
    run(step_abi, ring)
    if not args.no_graph:
        drain_exchange()
        torch.cuda.synchronize()
        for k in range(ring):
            replays[k] = sets[k]["step"].capture(frames[k], stage[k % NBUF]["G"] if train else None,
                                                 loss_weight=stage[k % NBUF]["G"] if train else None, stages=False)
        run(step_abi, ring)
    ms = timed(step_abi, n)
    for s in sets:
        s["step"].check_capacity()
    out = {"value": world * n / (ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": n, "ms_per_step": ms / n,
           "h2d_pinned_gbs_measured": round(h2d_gbs, 2),
           "h2d_gbs_implied": round(h2d * (n / (ms / 1e3)) / 1e9, 2),
           "pipeline": "inputs prefetched one step ahead on a copy stream; result read back one step late; "
                       "all copies inside the timed region" if train else
                       "pose prefetched one step ahead; frame converted to 8 bits on the device, copied to pinned memory on "
                       "a side stream and handed to host threads" + (" that JPEG-encode it (cv2)" if args.encode else "") +
                       "; the timed region ends when the last frame has been delivered",
           "api": "C ABI (include/sings_b200.h) driven by sings_b200.step.AvatarStep with preallocated buffers"
                  + ("" if args.no_graph else ", one CUDA graph per frame")}
    if train and not args.no_graph and not PROBE:
        # Variant: what a trainer does -- the step's TARGET image goes up as the dataset's uint8
        # (H, W, 3) (a quarter of the bytes of a float32 dL/dimage), and the image loss of the
        # reference (L1 + SSIM, HumanLoss.forward) and its gradient are computed on the device by
        # the fused loss kernels between forward and backward, inside the graph.  Reported beside
        # the float32-gradient headline, which is bound by the box's host link, not by the GPU.
        from sings_b200.losses import ImageLossBuffers
        upload["u8"] = True
        drain_exchange()
        torch.cuda.synchronize()
        loss_bufs = [ImageLossBuffers(H, W, dev) for _ in range(NBUF)]

        def loss_of(k):
            sb, lb, bgk = stage[k % NBUF], loss_bufs[k % NBUF], host[k]["bg"]

            def fn(img):
                dL = lb.run(img, sb["T8"], None, bgk)
                return lb.loss_value, dL
            return fn
        for k in range(ring):
            replays[k] = sets[k]["step"].capture(frames[k], None, loss_fn=loss_of(k), stages=False)
        run(step_abi, ring)
        ms8 = timed(step_abi, n)
        upload["u8"] = False
        out["u8_upload"] = {"value": world * n / (ms8 / 1e3), "unit": "frames/s", "steps": n, "ms_per_step": ms8 / n,
                            "h2d_bytes_per_step": h2d - 3 * H * W * 3, "d2h_bytes_per_step": 4,
                            "note": "same C-ABI loop; the target image is uploaded as uint8 (H, W, 3) and the reference's "
                                    "image loss (L1 + SSIM) and its gradient are computed on the device by the fused "
                                    "loss kernels (sgs_image_loss_fwd / _bwd) between forward and backward, inside the graph"}
    if params:
        nd = max(10, min(args.steps, 100))
        run(step_dropin, ring)   # checked mode: sizes the pair-list capacity for every avatar of the ring
        R.set_async(True)        # then no mid-step host sync: the overflow flag is examined at the next forward
        run(step_dropin, ring)
        msd = timed(step_dropin, nd)
        R.check_pending(block=True)
        R.set_async(False)
        out["dropin"] = {"value": world * nd / (msd / 1e3), "unit": "frames/s", "steps": nd, "ms_per_step": msd / nd,
                         "api": "sings_b200.deform.pose_to_A + deform_gaussians + "
                                "diff_gaussian_rasterization.GaussianRasterizer (torch.autograd)"}
        from sings_b200.fused import AvatarRenderer
        for p in params:
            renderers.append(AvatarRenderer(p["xyz"], p["rot"], p["scales"], p["opacity"], p["shs"], p["W"], p["rest"],
                                            p["parents"], p["inv_A"], H, W, D, sync_check=False))
            renderers[-1].step.L_cap = max(renderers[-1].step.L_cap, sets[0]["step"].L_cap)
            renderers[-1].step._alloc_scratch()
        run(step_fused, ring)
        for r in renderers:
            r.check()
        msf = timed(step_fused, nd)
        for r in renderers:
            r.check()
        out["fused_autograd"] = {"value": world * nd / (msf / 1e3), "unit": "frames/s", "steps": nd, "ms_per_step": msf / nd,
                                 "api": "sings_b200.fused.AvatarRenderer: the same computation as ONE torch.autograd "
                                        "Function over the fused kernels (opt-in, INTEGRATION.md 2b); loss = <image, G> and "
                                        ".backward() through autograd, gradients cloned out of the bucket"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="workload: BASELINE.json configs c1..c5 or the shipped configuration (default c2, the metric's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ab", action="store_true", help="skip the CUB / eager-torch comparators")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-dropin", action="store_true", help="skip the autograd drop-in variant of the e2e run")
    ap.add_argument("--encode", action="store_true",
                    help="forward-only configs: the e2e loop also JPEG-encodes every frame to gpurun_out/anim_rank*/ (cv2, "
                         "host threads), like the reference's animation loop")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if cfg.get("frames") and args.steps == 400:
        args.steps = cfg["frames"]
    if args.impl == "reference":
        reference_arm(args, cfg)
    else:
        gpu_arm(args, cfg)


if __name__ == "__main__":
    main()

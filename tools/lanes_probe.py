"""c4-shaped animation (120 frames, 200k Gaussians, 1080p, forward only) through sings_b200.animate.render_frames
with one AvatarStep and with two (two streams): frames/s of the user-facing call, images compared.
    python tools/lanes_probe.py > gpurun_out/lanes_probe.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from sings_b200 import synthetic as syn
from sings_b200.animate import render_frames
from sings_b200.step import AvatarStep, FrameInputs


def main():
    cfg = bench.CONFIGS["c4"]
    dev = torch.device("cuda", 0)
    N, H, W, D, J, F = cfg["N"], cfg["H"], cfg["W"], cfg["D"], cfg["J"], cfg["frames"]
    av, pose, transl, view, _, bg = bench.build_frame_inputs(cfg, 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    mk = lambda: AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs), t(av.lbs_weights),
                            t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), H, W, D)
    steps = [mk(), mk()]
    frames = [FrameInputs(pose=t(syn.random_pose(J, seed=500 + f)), transl=t(transl), viewmatrix=t(view.world_view_transform),
                          projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(bg),
                          tanfovx=view.tanfovx, tanfovy=view.tanfovy) for f in range(F)]
    out = torch.empty(F, 3, H, W, device=dev)
    res = {}
    imgs = {}
    for name, arg in (("one_lane", steps[0]), ("two_lanes", steps)):
        render_frames(arg, frames, out=out)                 # warm-up (capacities settle)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            render_frames(arg, frames, out=out)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[name] = {"frames_per_s": round(F / min(ts), 1), "ms_per_frame": round(min(ts) / F * 1e3, 4)}
        imgs[name] = out[::17].clone()
    res["speedup"] = round(res["two_lanes"]["frames_per_s"] / res["one_lane"]["frames_per_s"], 3)
    res["same_images"] = bool(torch.equal(imgs["one_lane"], imgs["two_lanes"]))
    res["note"] = f"render_frames, {F} frames, {N} Gaussians, {W}x{H}, forward only, host wall clock incl. launches and the copy into `out`"
    print(json.dumps(res))


if __name__ == "__main__":
    main()

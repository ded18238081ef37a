"""N-rank equivalence of the data-parallel step on the GPU under NCCL:
run as `python -m torch.distributed.run --nproc-per-node N tests/dp_nccl_check.py`.

Every rank renders its own turn-around view of the SAME avatar (replicated canonical Gaussians)
with AvatarStep, the buckets are exchanged with GradExchange (NCCL all-reduce SUM + MAX), and
every rank then checks, against all ranks' pre-exchange buckets collected with all_gather:
  reduced bucket == sum of the per-rank buckets (fp32 re-association tolerance),
  xyz_gradient_accum / denom == sums of the per-view statistics, max_radii2D == max over the views,
  and the reduced bucket is bit-identical on every rank.
Prints `DP_NCCL_CHECK PASS world=N` on rank 0.  tests/test_gpu_dp.py launches it when the box
has two GPUs; profiles/ keeps the output of the 2- and 8-GPU runs."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sings_b200 import dp, synthetic as syn          # noqa: E402
from sings_b200.step import AvatarStep, FrameInputs  # noqa: E402


def main():
    rank, local, world = dp.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, H, W, J, D = 30_000, 256, 256, 24, 3
    av = syn.make_avatar(N, J, seed=5)                       # the same avatar on every rank
    t = lambda a: torch.as_tensor(a, device=dev)
    params = [t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs), t(av.lbs_weights)]
    if rank != 0:
        for p in params:
            p.zero_()
    dp.broadcast_parameters(params, src=0)                   # replicas start from rank 0's parameters
    step = AvatarStep(*params[:5], params[5], t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), H, W, D)
    exch = dp.GradExchange(N, step.n_param_grads, dev)
    transl = syn.default_transl(H)
    G = torch.randn(3, H, W, device=dev, generator=torch.Generator(dev).manual_seed(7))     # same dL/dimage everywhere
    for it in range(2):                                      # two steps: the statistics accumulate across steps
        view = syn.make_view(H, W, yaw=2 * math.pi * (rank + world * it) / (2 * world), centre=(0.0, 0.0, float(transl[2])))
        fr = FrameInputs(pose=t(syn.random_pose(J, seed=11 + it)), transl=t(transl), viewmatrix=t(view.world_view_transform),
                         projmatrix=t(view.full_proj_transform), campos=t(view.camera_center),
                         bg=t(np.ones(3, np.float32)), tanfovx=view.tanfovx, tanfovy=view.tanfovy)
        step.forward(fr)
        step.backward(G)
        torch.cuda.synchronize()
        assert step.check_capacity() > 0
        local_bucket, local_radii = step.bucket.clone(), step.max_radii2D.clone()
        gathered = [torch.empty_like(local_bucket) for _ in range(world)]
        radii_all = [torch.empty_like(local_radii) for _ in range(world)]
        dist.all_gather(gathered, local_bucket)
        dist.all_gather(radii_all, local_radii)
        prev_accum, prev_denom, prev_max = exch.xyz_gradient_accum.clone(), exch.denom.clone(), exch.max_radii2D.clone()
        grads = exch.exchange(step.bucket, step.max_radii2D)          # SUM + MAX all-reduce, fold, clear step stats
        torch.cuda.synchronize()
        ref = torch.stack(gathered).double().sum(0)
        n = step.n_param_grads
        err = (grads.double() - ref[:n]).abs().max().item() / (ref[:n].abs().max().item() + 1e-30)
        assert err < 1e-6, f"reduced gradients differ from the sum of the per-rank buckets: {err}"
        assert torch.allclose(exch.xyz_gradient_accum.double(), prev_accum.double() + ref[n:n + N], rtol=1e-6, atol=1e-9)
        assert torch.equal(exch.denom, prev_denom + ref[n + N:].float())
        assert torch.equal(exch.max_radii2D, torch.maximum(prev_max, torch.stack(radii_all).max(0).values))
        assert float(step.bucket[n:].abs().max()) == 0.0 and float(step.max_radii2D.abs().max()) == 0.0
        # every rank holds the same reduced bits
        chk = torch.tensor([float(grads.double().sum()), float(grads.abs().double().sum())], device=dev, dtype=torch.float64)
        chks = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(chks, chk)
        assert all(torch.equal(c, chks[0]) for c in chks), "ranks disagree on the reduced bucket"
        assert not torch.equal(gathered[0], gathered[-1]) or world == 1      # the views really differ
    dist.barrier()
    if rank == 0:
        print(f"DP_NCCL_CHECK PASS world={world} N={N} bucket={step.bucket.numel() * 4 / 1e6:.1f} MB rel_err<1e-6", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

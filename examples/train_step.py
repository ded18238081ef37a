"""A complete optimisation loop on the device with the pieces of this repository: pose -> deform ->
rasterize (one autograd call), the reference's image loss (L1 + SSIM), scale-edge loss, L2Norm regulariser
and a Laplacian smoothing term, Adam on the Gaussian parameters.  Synthetic avatar and target (no SMPL assets needed):

    python examples/train_step.py [--steps 50] [--gaussians 50000] [--size 512]

What maps to what in SinGS: AvatarRenderer = sings_hybrid.py:390-428 + gs_renderer_single.py:48-101,
image_loss = HumanLoss.forward's image terms (loss.py:57-70), GaussiansEdgeLoss = loss_items.py:57-90,
L2Norm = loss_items.py:15-54, pcd_laplacian_smoothing = loss_items.py:205-214 (the region Laplacians of the
trainer need the SMPL vertex segmentation; the K-NN graph stands in for the mesh here).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sings_b200 import synthetic as syn
from sings_b200.fused import AvatarRenderer
from sings_b200.losses import GaussiansEdgeLoss, image_loss
from sings_b200.regularizers import L2Norm, build_edges, laplacian, pcd_laplacian_smoothing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--gaussians", type=int, default=50_000)
    ap.add_argument("--size", type=int, default=512)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    N, H, W, J, D = args.gaussians, args.size, args.size, 24, 3
    av = syn.make_avatar(N, J, seed=0, isotropic=True)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    view = syn.make_view(H, W)
    cam = (t(view.world_view_transform), t(view.full_proj_transform), t(view.camera_center), torch.ones(3, device=dev),
           view.tanfovx, view.tanfovy)
    pose, transl = t(syn.random_pose(J, seed=2)), t(syn.default_transl(H))

    def params(seed_noise):
        g = torch.Generator(dev).manual_seed(seed_noise)
        noise = lambda x, s: x + s * torch.randn(x.shape, device=dev, generator=g) if seed_noise else x
        return dict(xyz=noise(t(av.xyz_canon), 0.003).requires_grad_(True), scales=t(av.scales).requires_grad_(True),
                    opacity=t(av.opacity).requires_grad_(True), shs=noise(t(av.shs), 0.2).requires_grad_(True))

    def renderer(p):
        return AvatarRenderer(p["xyz"], None, p["scales"], p["opacity"], p["shs"], t(av.lbs_weights), t(av.rest),
                              torch.from_numpy(av.parents), t(av.inv_A_t2cano), H, W, D)
    # target: the clean avatar, rendered once and stored as a uint8 image like a dataset frame
    with torch.no_grad():
        target, _ = renderer(params(0))(pose, transl, *cam)
    gt_u8 = (target.clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous()
    # model: perturbed positions and colours
    p = params(1)
    r = renderer(p)
    opt = torch.optim.Adam([{"params": [p["xyz"]], "lr": 2e-4}, {"params": [p["shs"]], "lr": 5e-3},
                            {"params": [p["scales"], p["opacity"]], "lr": 1e-3}])
    edge = GaussiansEdgeLoss(K=9)
    l2 = L2Norm(lambda_xyz_offsets=0.001, lambda_scales_diff=0.005, max_scale_threshold=0.005, lambda_max_scale=0.01,
                min_opacity_threshold=0.2, lambda_min_opacity=0.001)             # human_complex.yaml:148-154
    xyz0 = p["xyz"].detach().clone()
    lap = laplacian(xyz0, build_edges(xyz0, 6))        # constant between densifications, like the trainer's operators
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(args.steps):
        image, radii = r(pose, transl, *cam)
        loss, items = image_loss(image, gt_u8, None, cam[3])
        loss = loss + 0.1 * edge({"xyz_canon": p["xyz"], "scales": p["scales"]})
        loss = loss + l2({"xyz_offsets": p["xyz"] - xyz0, "scales": p["scales"], "opacity": p["opacity"]})
        loss = loss + 0.01 * pcd_laplacian_smoothing(p["xyz"], lap)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        with torch.no_grad():
            p["scales"].clamp_(min=1e-5); p["opacity"].clamp_(1e-4, 1 - 1e-4)
        if it % 10 == 0 or it == args.steps - 1:
            print(f"step {it:4d}  loss {float(loss.detach()):.5f}  l1 {float(items['l1']):.5f}  ssim {float(items['ssim']):.5f}  "
                  f"visible {int((radii > 0).sum())}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{args.steps} optimisation steps of {N} Gaussians at {H}x{W}: {dt / args.steps * 1e3:.2f} ms per step")


if __name__ == "__main__":
    main()

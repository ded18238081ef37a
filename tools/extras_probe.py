"""Run the kernels beyond the per-frame path once each at BASELINE c2 sizes (ncu target + timings):
fused image loss, exact K-NN, tri-plane interpolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sings_b200 import synthetic as syn
from sings_b200.triplane import HexPlaneField
from sings_b200.losses import ImageLossBuffers, knn_points
dev = torch.device("cuda", 0)
g = torch.Generator(dev).manual_seed(0)
H = W = 1024
pred = torch.rand(3, H, W, device=dev, generator=g)
gt = (torch.rand(H, W, 3, device=dev, generator=g) * 255).to(torch.uint8)
lb = ImageLossBuffers(H, W, dev)
xyz = torch.as_tensor(syn.make_avatar(200_000, 24, seed=0).xyz_canon, device=dev).float().contiguous()
field = HexPlaneField({"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": 32, "resolution": [64, 64, 64],
                       "multires": [1, 2, 4]}, bounds=1.3, device=dev)
pts = xyz.clone().requires_grad_(True)
d_out = torch.randn(200_000, 96, device=dev)


def once():
    lb.run(pred, gt, None, torch.ones(3, device=dev))
    knn_points(xyz, 8)
    pts.grad = None
    (field(pts) * d_out).sum().backward()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


once(); once()
torch.cuda.synchronize()
print(f"image loss fwd+bwd 1024^2: {timed(lambda: lb.run(pred, gt, None, torch.ones(3, device=dev))):.1f} us")
print(f"knn K=8, 200k points (incl. scratch allocation): {timed(lambda: knn_points(xyz, 8)):.1f} us")
f = field(pts)
print(f"hexplane forward 200k x 96: {timed(lambda: field(pts)):.1f} us")
print(f"hexplane forward + backward: {timed(lambda: torch.autograd.grad(field(pts), [pts] + list(field.parameters())[1:], d_out)):.1f} us")

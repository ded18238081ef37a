"""Time the fused image-loss kernels alone (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sings_b200.losses import ImageLossBuffers
dev = torch.device("cuda", 0)
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = torch.Generator(dev).manual_seed(0)
pred = torch.rand(3, H, W, device=dev, generator=g)
gt = (torch.rand(H, W, 3, device=dev, generator=g) * 255).to(torch.uint8)
bg = torch.ones(3, device=dev)
lb = ImageLossBuffers(H, W, dev)
for _ in range(5):
    lb.run(pred, gt, None, bg)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    lb.run(pred, gt, None, bg)
b.record(); torch.cuda.synchronize()
print(f"image loss fwd+bwd {H}x{W}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per call, loss {float(lb.loss_value):.6f}")

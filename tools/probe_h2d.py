"""H2D / D2H bandwidth of pinned 12.6 MB buffers (the e2e arm's per-step upload) on this box."""
import torch, time
n = 3 * 1024 * 1024
h = torch.empty(n).pin_memory(); d = torch.empty(n, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"H2D pinned {n*4/1e6:.1f} MB: {ms:.3f} ms  {n*4/ms/1e6:.1f} GB/s")

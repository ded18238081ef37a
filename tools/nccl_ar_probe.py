"""All-reduce time of the gradient bucket (13.2 M floats = 52.8 MB) under torchrun -- NCCL knobs come from the environment."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13_200_000
x = torch.randn(n, device="cuda")
for _ in range(10):
    dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} bytes {n*4/1e6:.1f} MB: {t.item():.4f} ms  algbw {n*4/t.item()/1e6:.0f} GB/s  busbw {n*4/t.item()/1e6*2*(world-1)/world:.0f} GB/s  env " +
          " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_")), flush=True)
dist.destroy_process_group()

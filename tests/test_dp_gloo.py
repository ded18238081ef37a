"""CPU, world_size 2 (gloo): the data-parallel exchange of the hot path (sings_b200/dp.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sings_b200 import dp


def test_sharding_helpers():
    assert dp.shard_views(8, 1, 4) == [1, 5]
    assert dp.shard_views(3, 2, 4) == [2] and dp.shard_views(3, 3, 4) == []
    spans = [dp.shard_frames(120, r, 8) for r in range(8)]
    assert spans[0] == (0, 15) and spans[-1] == (105, 120)
    spans = [dp.shard_frames(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum(hi - lo for lo, hi in spans) == 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    N, n_param = 50, 7 * 50
    ex = dp.GradExchange(N, n_param, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    totals = torch.zeros(n_param)
    accum = torch.zeros(N)
    den = torch.zeros(N)
    mx = torch.zeros(N)
    for step in range(3):
        bucket = torch.rand(n_param + 2 * N, generator=g)
        bucket[n_param + N:] = (bucket[n_param + N:] > 0.5).float()        # denom increments are 0/1
        radii = torch.randint(0, 40, (N,), generator=g).float()
        # what every rank must end up with: recompute all ranks' contributions locally
        exp_sum = torch.zeros_like(bucket)
        exp_max = torch.zeros(N)
        for rr in range(world):
            gg = torch.Generator().manual_seed(100 + rr)
            for _ in range(step + 1):
                b = torch.rand(n_param + 2 * N, generator=gg)
                b[n_param + N:] = (b[n_param + N:] > 0.5).float()
                rd = torch.randint(0, 40, (N,), generator=gg).float()
            exp_sum += b
            exp_max = torch.maximum(exp_max, rd)
        finish = ex.exchange(bucket, radii, async_op=True, reset_step=False)    # fresh step buffers every iteration here
        grads = finish()
        assert torch.allclose(grads, exp_sum[:n_param], atol=1e-6)
        accum += exp_sum[n_param:n_param + N]
        den += exp_sum[n_param + N:]
        mx = torch.maximum(mx, exp_max)
        assert torch.allclose(ex.xyz_gradient_accum, accum, atol=1e-5)
        assert torch.allclose(ex.denom, den) and torch.equal(ex.max_radii2D, mx)
    # deferred MAX + reset_step: one collective per step; the radii become global at sync_max()
    ex2 = dp.GradExchange(N, n_param, "cpu", defer_max=True)
    local_mx = torch.zeros(N)
    all_mx = torch.zeros(N)
    for step in range(2):
        bucket = torch.full((n_param + 2 * N,), float(rank + 1))
        radii = torch.full((N,), float(10 * step + rank))
        radii[rank] = 99.0 + rank                       # an entry only this rank makes large
        local_mx = torch.maximum(local_mx, radii)
        for rr in range(world):
            rd = torch.full((N,), float(10 * step + rr))
            rd[rr] = 99.0 + rr
            all_mx = torch.maximum(all_mx, rd)
        grads = ex2.exchange(bucket, radii, async_op=True, reset_step=True)()
        assert torch.equal(grads, torch.full((n_param,), float(sum(range(1, world + 1)))))
        assert torch.equal(ex2.max_radii2D, local_mx)                  # still local
        assert float(bucket[n_param:].abs().max()) == 0.0 and float(radii.abs().max()) == 0.0   # step buffers cleared
    assert torch.equal(ex2.sync_max(), all_mx)
    assert torch.equal(ex2.denom, torch.full((N,), 2.0 * sum(range(1, world + 1))))
    t = torch.full((4,), float(rank))
    dp.broadcast_parameters([t], src=1)
    assert torch.equal(t, torch.ones(4))
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_exchange_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == [0, 1]


def test_single_process_exchange_is_local_accumulation():
    ex = dp.GradExchange(4, 8, "cpu")
    b = torch.arange(16, dtype=torch.float32)
    b0 = b.clone()
    g = ex.exchange(b, torch.tensor([1.0, 5.0, 0.0, 2.0]), reset_step=False)
    assert torch.equal(g, b[:8]) and torch.equal(ex.xyz_gradient_accum, b[8:12])
    ex.exchange(b, torch.tensor([3.0, 1.0, 0.0, 2.0]), reset_step=False)
    assert torch.equal(ex.max_radii2D, torch.tensor([3.0, 5.0, 0.0, 2.0]))
    assert torch.equal(ex.denom, 2 * b[12:16])
    # default (reset_step=True): the step's statistics are cleared after the fold, so an AvatarStep that
    # keeps accumulating into the same slices is not counted twice
    ex2 = dp.GradExchange(4, 8, "cpu")
    r = torch.tensor([1.0, 5.0, 0.0, 2.0])
    ex2.exchange(b, r)
    assert float(b[8:].abs().max()) == 0.0 and float(r.abs().max()) == 0.0 and torch.equal(b[:8], b0[:8])
    ex2.exchange(b, r)                                   # nothing new was accumulated: nothing is added
    assert torch.equal(ex2.xyz_gradient_accum, b0[8:12]) and torch.equal(ex2.denom, b0[12:16])

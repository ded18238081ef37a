"""Data-parallel step under NCCL on real GPUs (needs two): launches tests/dp_nccl_check.py with torchrun."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_nccl_step_equals_sum_of_per_rank_steps():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577",
                          os.path.join(ROOT, "tests", "dp_nccl_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DP_NCCL_CHECK PASS world=2" in out.stdout

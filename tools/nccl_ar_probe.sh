#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
run() { env "$@" $TR tools/nccl_ar_probe.py 2>/dev/null | grep world; }
run NCCL_MAX_CTAS=32
run NCCL_MAX_CTAS=64
run NCCL_DEBUG=WARN
run NCCL_ALGO=NVLS
run NCCL_ALGO=Ring
run NCCL_ALGO=Tree
run NCCL_MIN_CTAS=32 NCCL_MAX_CTAS=64 NCCL_ALGO=NVLS

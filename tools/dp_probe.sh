#!/bin/bash
# data-parallel probe (run under gpurun --gpus N): bench at N ranks for several NCCL CTA caps
N=${1:-2}; shift
for c in "$@"; do
  echo "=== N=$N NCCL_MAX_CTAS=$c"
  SGS_NCCL_MAX_CTAS=$c python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --no-dropin --no-ab 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sync value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), '| pipelined', round(d['dp']['pipelined']['value'],1), 'ms', round(d['dp']['pipelined']['ms_per_step'],4), '| e2e', round(d['e2e']['value'],1), '| bucket MB', d['dp']['bucket_mb'])
"
done

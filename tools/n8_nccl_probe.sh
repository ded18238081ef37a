# N-GPU check of the data-parallel exchange defaults (run under gpurun --gpus N)
N=${1:-8}
run() { echo "=== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), {k: round(v['ms']*1000,1) for k,v in d['roofline']['stages'].items()})"; }
run "SGS_NCCL_MAX_CTAS=8"
run "SGS_NCCL_MAX_CTAS=12"
run "SGS_NCCL_MAX_CTAS=16"
run "SGS_NCCL_MAX_CTAS=24"

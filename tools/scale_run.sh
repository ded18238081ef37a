#!/bin/bash
# multi-GPU evidence (run under gpurun --gpus N): NCCL equivalence check + bench lines at N ranks
N=${1:-2}; TAG=${2:-scale}; shift; shift
CFGS=${@:-c2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR tests/dp_nccl_check.py 2>&1 | grep -E "DP_NCCL_CHECK|Error|error" | head -5 | tee gpurun_out/${TAG}_n${N}_nccl_check.txt
for c in $CFGS; do
  $TR bench.py --gpus $N --config $c --steps 200 --warmup 10 --no-cpu --no-dropin --no-ab > gpurun_out/${TAG}_n${N}_$c.json 2> gpurun_out/${TAG}_n${N}_$c.err || tail -5 gpurun_out/${TAG}_n${N}_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_n${N}_$c.json").read().strip().splitlines()[-1])
    dp = d.get("dp") or {}
    print("$c N=$N value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "| pipelined", dp.get("pipelined", {}).get("value"), "| e2e", round(d["e2e"]["value"], 1), "| clocks", d["clocks"])
except Exception as e:
    print("$c FAILED", e)
PY
done

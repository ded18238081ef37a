TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581"
timeout 120 $TR tests/dp_nccl_check.py 2>&1 | grep -E "DP_NCCL_CHECK" | head -2 | tee gpurun_out/r02_d_n4_nccl_check.txt
timeout 240 $TR bench.py --gpus 4 --steps 200 --warmup 10 --no-dropin > gpurun_out/r02_d_n4_c2.json 2> gpurun_out/r02_d_n4_c2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_d_n4_c2.json").read().strip().splitlines()[-1])
print("N=4 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "| pipelined", d["dp"]["pipelined"]["value"], "| e2e", round(d["e2e"]["value"], 1), d["clocks"])
PY

#!/bin/bash
# round-2 GPU check: parity tests, short bench (stage table), launch list
TAG=${1:-r2}
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${TAG}_tests.txt
cat gpurun_out/${TAG}_tests.txt | tail -3
show='
import json,sys
d=json.loads(sys.stdin.read())
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "clk", d["clocks"])
for k,v in d["roofline"]["stages"].items(): print(" ", k, v)
print(" hot", d["roofline"]["lbs_preprocess_sort"])
'
python bench.py --steps 200 --warmup 10 --no-cpu --no-dropin 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
python -c "$show" < gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -30

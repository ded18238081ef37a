#!/bin/bash
# every north-star workload once on 1 GPU (run under gpurun): gpurun_out/<tag>_<config>.json + a summary
TAG=${1:-cfg}; shift
CFGS=${@:-c1 c3 c4 c5 shipped}
for c in $CFGS; do
  python bench.py --config $c --steps 100 --warmup 5 --no-cpu --no-dropin --no-ab > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err || tail -5 gpurun_out/${TAG}_$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$c.json"))
    r = d["roofline"]
    print("$c", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "L", int(r["pairs_L"]),
          "hot frac", r["lbs_preprocess_sort"]["frac"], "| " + " ".join(f"{k}={v['ms']*1000:.0f}us/{v['frac']:.2f}" for k, v in r["stages"].items()))
except Exception as e:
    print("$c FAILED", e)
PY
done

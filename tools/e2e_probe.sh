#!/bin/bash
# where does the e2e arm's time go?  pinned H2D bandwidth of this box, then the C-ABI e2e loop with
# (a) everything, (b) no dL/dimage upload, (c) no lagged loss read, (d) neither
python tools/probe_h2d.py
for p in "" nog nosync nog,nosync; do
  SGS_E2E_PROBE=$p python bench.py --steps 300 --warmup 10 --no-cpu --no-dropin 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('probe=[$p] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],4))
"
done

#!/bin/bash
# A/B helper: short bench (no e2e, no cpu) printing the stage table; env vars select variants
python bench.py --steps 200 --warmup 10 --no-cpu --no-e2e --no-ab 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), ' '.join(f\"{k}={v['ms']*1000:.1f}\" for k,v in d['roofline']['stages'].items()))
"

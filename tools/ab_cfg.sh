#!/bin/bash
# A/B helper: short bench of one config (no e2e, no cpu, no ab) printing the stage table
CFG=${1:-c2}
python bench.py --config $CFG --steps ${2:-100} --warmup 10 --no-cpu --no-e2e --no-ab 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$CFG value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), ' '.join(f\"{k}={v['ms']*1000:.1f}\" for k,v in d['roofline']['stages'].items()))
"

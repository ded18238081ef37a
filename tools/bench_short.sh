#!/bin/bash
# quick GPU check used during kernel iteration: parity tests, then a short bench with the stage table
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 200 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'dropin', round(d['e2e'].get('dropin',{}).get('value',0),1), 'clk', d['clocks'])
for k,v in d['roofline']['stages'].items(): print(' ', k, v)
print(' hot', d['roofline']['lbs_preprocess_sort'])
"

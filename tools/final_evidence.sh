#!/bin/bash
# Evidence of one version on one GPU (run under gpurun): GPU tests, the default bench line, the
# reference arm, the ncu launch list + full-set capture.   usage: bash tools/final_evidence.sh <tag>
TAG=${1:-vX}
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/${TAG}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
bash tools/profile.sh ${TAG} > /dev/null 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_compact.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_ncu_compact.txt
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_shares.txt
cat gpurun_out/${TAG}_tests.txt; tail -1 gpurun_out/${TAG}_smoke.txt; tail -c 600 gpurun_out/${TAG}_bench.json; echo; head -14 gpurun_out/${TAG}_launch_shares.txt

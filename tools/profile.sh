#!/bin/bash
# ncu evidence for one version: launch list (times) + full-set capture of steady-state steps
# usage: bash tools/profile.sh <tag> [config]     (run under gpurun; outputs in gpurun_out/)
TAG=${1:-prof}; CFG=${2:-c2}
CMD="python bench.py --config $CFG --steps 8 --warmup 4 --no-e2e --no-cpu --no-ab"
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 64 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -s 120 -c 17 -f -o gpurun_out/${TAG} $CMD \
    > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out | grep ${TAG}

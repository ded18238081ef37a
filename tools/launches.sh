#!/bin/bash
# ncu launch list only (per-kernel device times of two steady-state steps); usage: bash tools/launches.sh <tag>
TAG=${1:-l}
ncu --metrics gpu__time_duration.sum --clock-control none -s 96 -c 64 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv

#!/bin/bash
# A/B sweep on the GPU box: one line per variant in the file given as $1:
#     name | make arguments (EXTRA="-D..." FMAD_<file>=true ...) | [test] [ncu]
# Every variant rebuilds the library, optionally runs the GPU parity tests, then the short bench.
# usage (under gpurun): bash tools/sweep.sh tools/sweep_variants.txt > gpurun_out/sweep.txt 2>&1
while IFS='|' read -r name margs dotest; do
  name=$(echo $name | tr -d " +")
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  echo "=== $name :: $margs"
  if ! eval make -s -B -j16 -C sings_b200/csrc $margs > /tmp/make.log 2>&1; then tail -5 /tmp/make.log; continue; fi
  case "$dotest" in *test*) python -m pytest tests -x -q -m gpu 2>&1 | tail -3;; esac
  bash tools/ab.sh
  case "$dotest" in *ncu*) bash tools/launches.sh "sw_$name" | head -16;; esac
done < "$1"
# leave the default build behind
make -s -B -j16 -C sings_b200/csrc > /dev/null 2>&1

#!/bin/bash
# bench every variant library under sings_b200/lib/variants (run under gpurun)
for f in sings_b200/lib/variants/*.so; do
  echo "=== $(basename $f .so)"
  SGS_LIB_PATH=$PWD/$f bash tools/ab.sh
done

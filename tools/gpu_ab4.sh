#!/bin/bash
TAG=${1:-ab4}
OUT=gpurun_out
mkdir -p $OUT
for r in "0,16" "4,16" "8,16" "0,4" "4,8" "12,16"; do
  echo "levels $r"; MON_DEBUG_SCATTER_LEVELS=$r python tools/timeline.py 2>&1 | sed -n 2,3p | cut -c1-220
done | tee $OUT/${TAG}_scatter_levels.txt

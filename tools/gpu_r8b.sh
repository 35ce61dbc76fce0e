#!/bin/bash
TAG=${1:-r8b}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline.txt
MON_SWEEP_EDGE_LATE=1 timeout 300 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_late_edge.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline_late_edge.txt
timeout 300 python tools/scatter_crossover.py > $OUT/${TAG}_scatter_crossover.txt 2>&1; tail -1 $OUT/${TAG}_scatter_crossover.txt
for e in 0 1; do MON_SWEEP_EDGE_LATE=$e timeout 200 python tools/quick_rate.py; done

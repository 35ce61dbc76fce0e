#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | grep "iter 1[78]\|iter 41[23]\|mean"
MON_PDL_MASK=6 timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_nopdl_o.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline_nopdl_o.txt | grep "iter 1[78]\|iter 41[23]\|mean"
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 300; echo

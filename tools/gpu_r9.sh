#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/gpu_r9.sh <tag>': GPU tests + device timelines in both phases + driver-shaped bench line
TAG=${1:-r9}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 700 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -4 $OUT/${TAG}_pytest_gpu.log | head -2
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | grep "iter 1[78]\|iter 41[23]\|mean"
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
head -c 400 $OUT/${TAG}_bench_20_5.json; echo
python tools/quick_rate.py 2>&1 | tail -1
python tools/quick_rate.py --rays 1024 --hidden-layers 2 2>&1 | tail -1
timeout 200 python tools/tc_phase_times.py > $OUT/${TAG}_mlp_phase_times.txt 2>&1; tail -42 $OUT/${TAG}_mlp_phase_times.txt | cut -c1-150

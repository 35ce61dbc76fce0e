#!/bin/bash
TAG=${1:-r5k}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline.txt
timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms|wall|rc " $OUT/${TAG}_facade_runs.txt | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 300; echo

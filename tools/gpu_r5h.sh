#!/bin/bash
# unified scatter kernel (one launch, two paths): full GPU parity + stage times + timelines of both phases + bench
TAG=${1:-r5h}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 200 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times.txt 2>&1
cat $OUT/${TAG}_stage_times.txt
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | head -30
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 300; echo

#!/bin/bash
TAG=${1:-r5n}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log
MON_INGEST_TRACE=1 timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms|wall|rc |mon ingest" $OUT/${TAG}_facade_runs.txt | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 300; echo
timeout 300 python bench.py --no-secondary > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json | head -c 300; echo

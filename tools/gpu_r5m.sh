#!/bin/bash
TAG=${1:-r5m}
OUT=gpurun_out
mkdir -p $OUT
MON_INGEST_TRACE=1 timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms|wall|rc |mon ingest" $OUT/${TAG}_facade_runs.txt | cut -c1-300

#!/bin/bash
TAG=${1:-r5l}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "resident or stage_by_stage" ) > $OUT/${TAG}_pytest_parity.log 2>&1
tail -2 $OUT/${TAG}_pytest_parity.log
timeout 200 python tools/timeline.py --at 5 > $OUT/${TAG}_timeline.txt 2>&1
grep "iter 1[78]\|mean" $OUT/${TAG}_timeline.txt
grep -A3 "resident scatter per CTA" $OUT/${TAG}_timeline.txt | tail -1 | cut -c1-600
MON_INGEST_TRACE=1 timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms|wall|rc |mon ingest" $OUT/${TAG}_facade_runs.txt | cut -c1-300

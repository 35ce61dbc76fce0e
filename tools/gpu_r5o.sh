#!/bin/bash
TAG=${1:-r5o}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/graph_cost_probe.py > $OUT/${TAG}_graph_cost.json 2>&1; cat $OUT/${TAG}_graph_cost.json
for rep in 1 2; do
MON_INGEST_TRACE=1 timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs_$rep.txt
grep -E "ingest_ms|mon ingest" $OUT/${TAG}_facade_runs_$rep.txt | cut -c1-300
done

#!/bin/bash
TAG=${1:-r5j}
OUT=gpurun_out
mkdir -p $OUT
MON_EXTRA_NVCC_FLAGS= ; 
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline.txt
MON_PDL_MASK=14 timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_pdl_scatter.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline_pdl_scatter.txt
timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms|wall|rc " $OUT/${TAG}_facade_runs.txt | cut -c1-300
timeout 120 python tools/ingest_probe.py > $OUT/${TAG}_ingest_probe.txt 2>&1; cat $OUT/${TAG}_ingest_probe.txt

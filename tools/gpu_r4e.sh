#!/bin/bash
TAG=${1:-r4e}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q ) > $OUT/${TAG}_pytest_parity.log 2>&1
tail -3 $OUT/${TAG}_pytest_parity.log
( MON_SCATTER_SMEM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q ) > $OUT/${TAG}_pytest_parity_smem.log 2>&1
tail -3 $OUT/${TAG}_pytest_parity_smem.log
timeout 200 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times.txt 2>&1
cat $OUT/${TAG}_stage_times.txt
MON_SCATTER_SMEM=1 timeout 200 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times_smem.txt 2>&1
cat $OUT/${TAG}_stage_times_smem.txt
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | head -20
MON_SCATTER_SMEM=1 timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_smem.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline_smem.txt | head -20

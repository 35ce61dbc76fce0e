#!/bin/bash
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
{
  MON_PIPE=0 python tools/quick_rate.py
  MON_PIPE=4 MON_SPLIT_LEVEL=8 python tools/quick_rate.py
  MON_PIPE=4 MON_SPLIT_LEVEL=4 python tools/quick_rate.py
  MON_PIPE=4 MON_SPLIT_LEVEL=12 python tools/quick_rate.py
} 2>&1 | tee $OUT/${TAG}_rates.txt
{ MON_PIPE=4 MON_SPLIT_LEVEL=8 python tools/timeline.py; MON_PIPE=4 MON_SPLIT_LEVEL=4 python tools/timeline.py; } > $OUT/${TAG}_timeline.txt 2>&1
grep -A2 graph_us $OUT/${TAG}_timeline.txt | cut -c1-420
( time MON_PIPE=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > $OUT/${TAG}_pytest_pipe4.log 2>&1
tail -4 $OUT/${TAG}_pytest_pipe4.log | head -1

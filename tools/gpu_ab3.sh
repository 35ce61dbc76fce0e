#!/bin/bash
TAG=${1:-ab3}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
grep -E "passed|failed" $OUT/${TAG}_pytest.log
{
  python tools/quick_rate.py
  echo "MON_OPT_SPEC=0"; MON_OPT_SPEC=0 python tools/quick_rate.py
  python tools/quick_rate.py --rays 1024 --hidden-layers 2
  echo "MON_OPT_SPEC=0"; MON_OPT_SPEC=0 python tools/quick_rate.py --rays 1024 --hidden-layers 2
} 2>&1 | tee $OUT/${TAG}_rates.txt
{ python tools/timeline.py; python tools/timeline.py --rays 1024 --hidden-layers 2; } > $OUT/${TAG}_timeline.txt 2>&1
grep -A2 graph_us $OUT/${TAG}_timeline.txt | cut -c1-300

#!/bin/bash
TAG=${1:-ab2}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log | head -1
{
  python tools/quick_rate.py
  for u in 1 3 4; do echo "MON_ENC_UNROLL=$u"; MON_ENC_UNROLL=$u python tools/quick_rate.py; done
  for c in 512 444 296; do echo "MON_MLP_CTAS=$c"; MON_MLP_CTAS=$c python tools/quick_rate.py; done
} 2>&1 | tee $OUT/${TAG}_rates.txt
{ python tools/timeline.py; MON_MLP_CTAS=512 python tools/timeline.py; python tools/timeline.py --rays 1024 --hidden-layers 2; } > $OUT/${TAG}_timeline.txt 2>&1
grep -A2 graph_us $OUT/${TAG}_timeline.txt | cut -c1-300

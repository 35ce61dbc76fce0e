#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/gpu_pipe_ab.sh <tag>': GPU parity tests + iteration-graph A/B with device timelines
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
{
  MON_PIPE=0 python tools/quick_rate.py
  for parts in 1 2 4; do MON_PIPE=1 MON_PIPE_ENC_PARTS=$parts python tools/quick_rate.py; done
  MON_PIPE=2 python tools/quick_rate.py
  MON_PIPE=3 python tools/quick_rate.py
  MON_PIPE=0 python tools/quick_rate.py --rays 1024 --hidden-layers 2
  MON_PIPE=1 python tools/quick_rate.py --rays 1024 --hidden-layers 2
  MON_PIPE=1 MON_PIPE_ENC_PARTS=1 python tools/quick_rate.py --rays 1024 --hidden-layers 2
} 2>&1 | tee $OUT/${TAG}_rates.txt
{
  MON_PIPE=0 python tools/timeline.py
  for parts in 1 2 4; do MON_PIPE=1 MON_PIPE_ENC_PARTS=$parts python tools/timeline.py; done
  MON_PIPE=2 python tools/timeline.py
} > $OUT/${TAG}_timeline.txt 2>&1
grep -E "graph_us|mean_period" $OUT/${TAG}_timeline.txt

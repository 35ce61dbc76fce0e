#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/gpu_quick.sh <tag>': GPU parity tests + steady-state rate + device timeline of the default graph
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log | head -1
{ python tools/quick_rate.py; python tools/quick_rate.py --rays 1024 --hidden-layers 2; } 2>&1 | tee $OUT/${TAG}_rates.txt
python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1
head -3 $OUT/${TAG}_timeline.txt | cut -c1-400

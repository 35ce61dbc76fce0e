#!/bin/bash
# gpurun --timeout 900 -- 'bash tools/gpu_ab.sh <tag> VAR=a,b [quick_rate args]': parity tests, then steady-state rate and device
# timeline for every value of one A/B environment switch (each value in its own process)
TAG=$1; shift
VAR=${1%%=*}; VALS=${1#*=}; shift
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
grep -E "passed|failed" $OUT/${TAG}_pytest.log
for v in ${VALS//,/ }; do
  echo "$VAR=$v"
  env $VAR=$v python tools/quick_rate.py "$@"
  env $VAR=$v python tools/quick_rate.py --rays 1024 --hidden-layers 2
  env $VAR=$v python tools/timeline.py "$@" 2>&1 | grep -E "^iter 42[01]|mean_period" | cut -c1-260
done 2>&1 | tee $OUT/${TAG}_ab.txt

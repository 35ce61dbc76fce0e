#!/bin/bash
# A/B of the graph structure: batch + points of the next iteration beside the hash encode (MON_EARLY_FORK, default on) and the
# scatter fused into the MLP kernel in steady state (MON_SCATTER_FUSED: -1 never, default by the live count)
TAG=${1:-r9c}
OUT=gpurun_out
mkdir -p $OUT
flt() { grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $1 | grep "iter 1[78]\|iter 41[23]\|mean"; }
for cfg in "default:" "nofuse:MON_SCATTER_FUSED=-1" "old:MON_EARLY_FORK=0 MON_SCATTER_FUSED=-1" "latefork_fused:MON_EARLY_FORK=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== $name ($envs)"
  env $envs timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_${name}.txt 2>&1
  flt $OUT/${TAG}_timeline_${name}.txt
  env $envs python tools/quick_rate.py 2>&1 | tail -1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
head -c 300 $OUT/${TAG}_bench_20_5.json; echo
python -c "
import json; d=json.loads(open('$OUT/${TAG}_bench_20_5.json').read().strip().splitlines()[-1]); print('e2e', d['e2e'])"
python tools/quick_rate.py --rays 1024 --hidden-layers 2 2>&1 | tail -1
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or u16 or graph or resident" ) 2>&1 | tail -3

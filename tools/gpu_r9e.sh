#!/bin/bash
# A/B: speculative two-quad optimizer sweep (MON_OPT_SPEC=1), scatter carve-out (MON_SCATTER_CARVEOUT)
TAG=${1:-r9e}
OUT=gpurun_out
mkdir -p $OUT
flt() { grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $1 | grep "iter 1[78]\|iter 41[23]\|mean"; }
for cfg in "base:MON_X=0" "optspec:MON_OPT_SPEC=1" "carve100:MON_SCATTER_CARVEOUT=100" "carve100_pdl:MON_SCATTER_CARVEOUT=100 MON_PDL_MASK=30"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== $name ($envs)"
  env $envs timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline_${name}.txt 2>&1
  flt $OUT/${TAG}_timeline_${name}.txt
  env $envs python tools/quick_rate.py 2>&1 | tail -1
done
( MON_OPT_SPEC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "optimizer or stage_by_stage or fused or graph" ) 2>&1 | tail -3
MON_OPT_SPEC=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5_optspec.json 2> $OUT/${TAG}_bench_20_5.err
head -c 260 $OUT/${TAG}_bench_20_5_optspec.json; echo

#!/bin/bash
# resident scatter v3 (parity x feature jobs, cost-balanced split, 140 CTAs): parity + timeline + ncu of the dense phase
TAG=${1:-r5c}
OUT=gpurun_out
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > $OUT/${TAG}_pytest_parity.log 2>&1
tail -5 $OUT/${TAG}_pytest_parity.log
timeout 200 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times.txt 2>&1
cat $OUT/${TAG}_stage_times.txt
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | head -30
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node \
    -k regex:'k_scatter_resident|k_optimizer_sweep' --launch-skip 20 --launch-count 2 -f -o $OUT/${TAG}_scatter_dense python tools/ncu_target.py --warm 5 > $OUT/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_scatter_dense.ncu-rep k_scatter_resident 30 > $OUT/${TAG}_ncu_scatter_dense.txt 2>&1
head -60 $OUT/${TAG}_ncu_scatter_dense.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 600; echo

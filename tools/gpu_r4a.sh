#!/bin/bash
# round 2, first GPU call: does the fused scatter + Adam path work, and what does it cost
TAG=${1:-r4a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > $OUT/${TAG}_pytest_parity.log 2>&1
tail -5 $OUT/${TAG}_pytest_parity.log
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py ) > $OUT/${TAG}_memcheck.log 2>&1
tail -4 $OUT/${TAG}_memcheck.log
( time timeout 900 python -m pytest tests/test_gpu_vs_reference_live.py tests/test_gpu_bench_config.py -q -k "not psnr" ) > $OUT/${TAG}_pytest_bench_shape.log 2>&1
tail -8 $OUT/${TAG}_pytest_bench_shape.log
timeout 200 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times.txt 2>&1
cat $OUT/${TAG}_stage_times.txt
MON_SCATTER_SMEM=1 timeout 200 python tools/stage_times.py --at 0,500 > $OUT/${TAG}_stage_times_smem.txt 2>&1
cat $OUT/${TAG}_stage_times_smem.txt
timeout 200 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
head -30 $OUT/${TAG}_timeline.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --cpu-seconds 3 > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
head -c 2500 $OUT/${TAG}_bench_20_5.json; echo; tail -3 $OUT/${TAG}_bench_20_5.err
timeout 300 python bench.py --no-secondary --cpu-seconds 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
head -c 1200 $OUT/${TAG}_bench.json; echo; tail -3 $OUT/${TAG}_bench.err

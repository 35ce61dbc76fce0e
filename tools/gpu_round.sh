#!/bin/bash
# One gpurun call's worth of validation + measurement (B200): GPU parity tests, both bench arms, the secondary
# configurations, the iteration-graph A/B (MON_PIPE), PSNR vs the reference Core, and the ncu launch list.
# Everything lands in gpurun_out/.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [tag] [nopsnr]'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
if ! grep -q " passed" $OUT/${TAG}_pytest.log || grep -q "failed" $OUT/${TAG}_pytest.log; then
  ( time MON_PIPE=0 timeout 900 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest_pipe0.log 2>&1
  tail -3 $OUT/${TAG}_pytest_pipe0.log
fi
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
# iteration-graph A/B: chain graph (default) vs the opt-in level-pipelined graph
for m in 0 1; do MON_PIPE=$m python tools/quick_rate.py; done 2>&1 | tee $OUT/${TAG}_pipe_ab.txt
timeout 300 python bench.py --rays 1024 --hidden-layers 2 --cpu-seconds 2 > $OUT/${TAG}_bench_r1024_nh2.json 2>> $OUT/${TAG}_bench.err
timeout 300 python bench.py --objects 4 --cpu-seconds 2 > $OUT/${TAG}_bench_4obj.json 2>> $OUT/${TAG}_bench.err
timeout 120 python tools/stage_times.py > $OUT/${TAG}_stage_times.txt 2>&1
if [ "$2" != "nopsnr" ]; then
  timeout 600 python tools/psnr_compare.py --iters 2000 --objects 2 --seeds 2 > $OUT/${TAG}_psnr.jsonl 2> $OUT/${TAG}_psnr.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 1400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 50 --warmup 250 --frames 8 --cpu-seconds 1 > $OUT/${TAG}_ncu_bench.log 2>&1
# one steady-state iteration under ncu --set full (cold caches, serialised: shares and per-kernel counters, not times)
timeout 600 ncu --set full --clock-control none --import-source on --graph-profiling node \
    -k regex:'k_encode_forward|k_mlp_train_tc|k_encode_backward|k_optimizer_sweep|k_generate_batch|k_sample_points' \
    --launch-skip 2400 --launch-count 6 -f -o $OUT/${TAG}_full python tools/ncu_target.py > $OUT/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_full.ncu-rep > $OUT/${TAG}_ncu_kernels.txt 2>&1
python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1
head -c 1800 $OUT/${TAG}_bench.json; echo
head -c 600 $OUT/${TAG}_bench_ref.json; echo
tail -1 $OUT/${TAG}_psnr.jsonl 2>/dev/null
tail -4 $OUT/${TAG}_stage_times.txt

#!/bin/bash
# Round-2 evidence in one gpurun call (B200 x1): GPU tests, both bench arms at the driver's shape and at the default, ncu launch
# list of the driver-shaped command, ncu --set full of one fresh-object iteration and one steady-state iteration, device timelines,
# stage times incl. the scatter paths' crossover, the C++ facade runs (configs 3 and 5).  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r9'
TAG=${1:-r9}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/${TAG}_pytest_gpu.log 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref_20_5.json 2> $OUT/${TAG}_bench_ref.err
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench.err
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err
python - <<PY
import json
for f in ("${OUT}/${TAG}_bench_ref_20_5.json", "${OUT}/${TAG}_bench_20_5.json", "${OUT}/${TAG}_bench.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), r.get("kernel"), r.get("frac") and round(r["frac"], 3), {k: round(v["ms"] * 1e3, 1) for k, v in (r.get("stages") or {}).items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python tools/stage_times.py --at 0,50,500 > $OUT/${TAG}_stage_times.txt 2>&1
timeout 300 python tools/scatter_crossover.py > $OUT/${TAG}_scatter_crossover.txt 2>&1; tail -1 $OUT/${TAG}_scatter_crossover.txt
timeout 300 python tools/timeline.py --at 5,400 > $OUT/${TAG}_timeline.txt 2>&1
grep "iter 1[78]\|iter 41[23]\|mean" $OUT/${TAG}_timeline.txt
# launch list of the driver-shaped command (cold caches, serialised: shares, not times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 400 --csv --log-file $OUT/${TAG}_launches_20_5.csv \
    python bench.py --steps 20 --warmup 5 --no-secondary --cpu-seconds 1 > $OUT/${TAG}_ncu_bench.log 2>&1
# one fresh-object iteration (iteration 8) and one steady-state iteration under ncu --set full
KERNELS='k_encode_forward|k_mlp_train_tc|k_scatter|k_optimizer_sweep|k_generate_batch|k_sample_points'
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:"$KERNELS" \
    --launch-skip 48 --launch-count 6 -f -o $OUT/${TAG}_full_fresh python tools/ncu_target.py --warm 5 > $OUT/${TAG}_ncu_fresh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:"$KERNELS" \
    --launch-skip 2400 --launch-count 9 -f -o $OUT/${TAG}_full_steady python tools/ncu_target.py > $OUT/${TAG}_ncu_steady.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_full_fresh.ncu-rep k_scatter 24 > $OUT/${TAG}_ncu_kernels_fresh_object.txt 2>&1
python tools/ncu_summary.py $OUT/${TAG}_full_steady.ncu-rep k_encode_forward 24 > $OUT/${TAG}_ncu_kernels_steady_state.txt 2>&1
python tools/ncu_summary.py --traffic $OUT/${TAG}_traffic.json scatter=$OUT/${TAG}_full_fresh.ncu-rep:k_scatter encode=$OUT/${TAG}_full_steady.ncu-rep:k_encode_forward \
    mlp_fused=$OUT/${TAG}_full_steady.ncu-rep:k_mlp_train_tc optimizer=$OUT/${TAG}_full_steady.ncu-rep:k_optimizer_sweep scatter_steady=$OUT/${TAG}_full_steady.ncu-rep:k_scatter optimizer_fresh=$OUT/${TAG}_full_fresh.ncu-rep:k_optimizer_sweep
grep -A14 "k_scatter" $OUT/${TAG}_ncu_kernels_fresh_object.txt | head -16
timeout 600 python tools/occupancy_report.py --seeds 2 > $OUT/${TAG}_occupancy_report.jsonl 2>&1; tail -2 $OUT/${TAG}_occupancy_report.jsonl
timeout 900 bash tools/gpu_facade_runs.sh > $OUT/${TAG}_facade.log 2>&1
cp $OUT/facade_runs.txt $OUT/${TAG}_facade_runs.txt
grep -E "ingest_ms_per|ingest_ms min|wall|rc " $OUT/${TAG}_facade_runs.txt | cut -c1-200

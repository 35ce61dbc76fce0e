#!/bin/bash
# 8-GPU check of the driver-shaped lines (one node): ours and the reference arm (8 objects on 8 GPUs from 8 host threads)
TAG=${1:-r7s}
OUT=gpurun_out
mkdir -p $OUT
N=${2:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_${N}gpu_20_5.json 2> $OUT/${TAG}_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_ref_${N}gpu_20_5.json 2> $OUT/${TAG}_bench_ref_${N}gpu.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench_${N}gpu_20_5.json", "$OUT/${TAG}_bench_ref_${N}gpu_20_5.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("seconds"), d["config"]["objects"], d["n_gpus"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err").replace("_20_5", "")).read()[-1500:])
PY

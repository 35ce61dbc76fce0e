#!/bin/bash
TAG=${1:-r5p}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python tools/timeline.py --at 5 > $OUT/${TAG}_timeline.txt 2>&1
grep "per-iteration\|mean" $OUT/${TAG}_timeline.txt | cut -c1-3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_2gpu_20_5.json 2> $OUT/${TAG}_bench_2gpu.err
head -c 1200 $OUT/${TAG}_bench_2gpu_20_5.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $OUT/${TAG}_bench_2gpu.json 2>> $OUT/${TAG}_bench_2gpu.err
python -c "
import json
for f in ('$OUT/${TAG}_bench_2gpu_20_5.json','$OUT/${TAG}_bench_2gpu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('h2d_bytes_per_step'))
"
tail -3 $OUT/${TAG}_bench_2gpu.err

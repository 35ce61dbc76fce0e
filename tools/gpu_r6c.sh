#!/bin/bash
TAG=${1:-r6c}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1200 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_parity.py tests/test_host_facade.py -m gpu -x -q -k "not psnr" ) > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_bench_2gpu_20_5.json 2> $OUT/${TAG}_bench_2gpu.err
python -c "
import json
for f in ('$OUT/${TAG}_bench_20_5.json', '$OUT/${TAG}_bench_2gpu_20_5.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('seconds'))
"
tail -2 $OUT/${TAG}_bench_2gpu.err
# 2-GPU facade: online replay with NVLink frame replication
mkdir -p /tmp/room && cd /tmp/room && python -c "import sys; sys.path.insert(0, '/root/repo'); from ro_map_b200 import synthetic as s; s.write_sequence(s.make_sequence(30, 4), 'room_synth')"
timeout 300 /root/repo/ro_map_b200/online_replay /root/repo/ro_map_b200/configs/base.json room_synth 1 500 4 out_online > on2.log 2>&1; grep -E "^object|ingest_ms_per|ingest_ms min" on2.log | cut -c1-200

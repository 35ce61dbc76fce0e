#!/bin/bash
# Round-end validation in one short gpurun call: GPU parity tests, smoke(), both bench arms, PSNR vs the reference Core.
#   gpurun --timeout 330 -- 'bash tools/gpu_final.sh [tag]'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 150 python -m pytest tests -m gpu -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 100 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 100 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 100 python tools/psnr_compare.py --iters 2000 --objects 2 --seeds 3 > $OUT/${TAG}_psnr.jsonl 2> $OUT/${TAG}_psnr.err
head -c 900 $OUT/${TAG}_bench.json; echo
head -c 300 $OUT/${TAG}_bench_ref.json; echo
tail -1 $OUT/${TAG}_psnr.jsonl

#!/bin/bash
# quick look: parity of the scatter paths + dense-phase timeline + driver-shaped bench line
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "resident or stage_by_stage or graph_path" ) > $OUT/${TAG}_pytest_parity.log 2>&1
tail -3 $OUT/${TAG}_pytest_parity.log
timeout 200 python tools/timeline.py --at 5 > $OUT/${TAG}_timeline.txt 2>&1
grep -v "^encode per-CTA\|^table resident\|^[0-9. ]*$" $OUT/${TAG}_timeline.txt | head -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary > $OUT/${TAG}_bench_20_5.json 2> $OUT/${TAG}_bench_20_5.err
cat $OUT/${TAG}_bench_20_5.json | head -c 300; echo

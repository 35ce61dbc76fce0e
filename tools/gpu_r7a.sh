#!/bin/bash
TAG=${1:-r7a}
OUT=gpurun_out
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_occupancy.py tests/test_gpu_parity.py -x -q ) > $OUT/${TAG}_pytest.log 2>&1
tail -15 $OUT/${TAG}_pytest.log
cat $OUT/occupancy_mode_test.json 2>/dev/null
timeout 600 python tools/occupancy_report.py > $OUT/${TAG}_occupancy_report.jsonl 2>&1; cat $OUT/${TAG}_occupancy_report.jsonl

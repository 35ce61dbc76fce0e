#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck initcheck; do
  ( time timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py ) > $OUT/r9_sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|Error|hazard" $OUT/r9_sanitize_$tool.log | head -8
done

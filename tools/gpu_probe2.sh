#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $OUT/dsmem_probe tools/dsmem_gather_probe.cu && timeout 120 $OUT/dsmem_probe > $OUT/r6b_dsmem_gather_probe.txt 2>&1
cat $OUT/r6b_dsmem_gather_probe.txt
timeout 120 python tools/h2d_probe.py > $OUT/r6b_h2d_probe.json 2>&1; cat $OUT/r6b_h2d_probe.json

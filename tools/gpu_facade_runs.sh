#!/bin/bash
# BASELINE.json configs 3 and 5 through the C++ drop-in layer (libMON.so): the headless OfflineNeRF driver (4 objects,
# 2000 iterations each = 4 Train_Steps of 500) and the online replay (NerfManagerOnline driven like the SLAM frontend).
OUT=${GRAFT_REPO_ROOT:-.}/gpurun_out
mkdir -p $OUT /tmp/room && cd /tmp/room
python -c "import sys; sys.path.insert(0, '/root/repo'); from ro_map_b200 import synthetic as s; s.write_sequence(s.make_sequence(30, 4), 'room_synth')"
BIN=/root/repo/ro_map_b200
{
  echo "== offline_nerf base.json room_synth 1 <steps=4> <objects=4> (config 3)"
  t0=$(date +%s.%N); $BIN/offline_nerf $BIN/configs/base.json room_synth 1 4 4 > off.log 2>&1; rc=$?; t1=$(date +%s.%N)
  grep -E "^object|train_time" off.log | head -30; echo "rc $rc wall $(python -c "print(round($t1 - $t0, 2))") s (incl. PNG decode of 30 keyframes, 4 x 2000 iterations, test view + 60-view video + mesh per object)"; tail -3 off.log
  echo "== online_replay base.json room_synth 1 <iters=500> <objects=4> (config 5)"
  t0=$(date +%s.%N); $BIN/online_replay $BIN/configs/base.json room_synth 1 500 4 out_online > on.log 2>&1; rc=$?; t1=$(date +%s.%N)
  grep -E "^object|ingest" on.log | cut -c1-400; echo "rc $rc wall $(python -c "print(round($t1 - $t0, 2))") s"; tail -3 on.log
  ls out_online/0 2>&1 | tr '\n' ' '; echo; ls out_online/0/video_img 2>/dev/null | wc -l
} > $OUT/facade_runs.txt 2>&1
cat $OUT/facade_runs.txt

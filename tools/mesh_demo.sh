#!/bin/bash
# trains the 4 objects of the default synthetic sequence through the headless OfflineNeRF and keeps the PLY meshes
set -e
out=${1:-gpurun_out/mesh_demo}
mkdir -p "$out"
python -c "from ro_map_b200 import synthetic as s; s.write_sequence(s.make_sequence(30, 4, H=400, W=400, K=(555.555, 555.555, 200.0, 200.0)), '/tmp/room_synth')"
cd "$out"
"$GRAFT_REPO_ROOT"/ro_map_b200/offline_nerf "$GRAFT_REPO_ROOT"/ro_map_b200/configs/base.json /tmp/room_synth 1 > log.txt 2>&1
tail -5 log.txt
ls -la output
cp /tmp/room_synth/obj_offline/*.txt . 2>/dev/null || true

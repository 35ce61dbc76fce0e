#!/usr/bin/env python
"""Target for `ncu --set full`: one object in steady state, then a few more graph iterations (the captured ones).
   ncu --set full --clock-control none --import-source on --graph-profiling node \
       -k regex:'k_encode_forward|k_mlp_train_tc|k_encode_backward|k_optimizer_sweep|k_generate_batch|k_sample_points' \
       --launch-skip 2400 --launch-count 6 -o gpurun_out/<tag> python tools/ncu_target.py"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--hidden-layers", type=int, default=1)
ap.add_argument("--warm", type=int, default=400)
a = ap.parse_args()
seq = syn.make_sequence(30, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
g = core.NerfObject(ds, core.default_config(rays_per_batch=a.rays, n_hidden_layers=a.hidden_layers), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
g.train(a.warm)      # 8 x 50-iteration graphs
print("loss", g.train(50))

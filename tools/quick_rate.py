#!/usr/bin/env python
"""Steady-state graph throughput of one object (product library), for A/B runs of MON_PIPE / MON_PIPE_ENC_PARTS /
MON_PDL_MASK in separate processes: python tools/quick_rate.py [--rays 4096] [--hidden-layers 1]"""
import argparse, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--hidden-layers", type=int, default=1)
ap.add_argument("--frames", type=int, default=30)
a = ap.parse_args()
seq = syn.make_sequence(a.frames, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
g = core.NerfObject(ds, core.default_config(rays_per_batch=a.rays, n_hidden_layers=a.hidden_layers), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
g.train(400)
rates = []
for _ in range(3):
    loss = g.train(500)
    rates.append(500.0 / (g.last_train_ms * 1e-3))
print(json.dumps({"MON_PIPE": os.environ.get("MON_PIPE", "default"), "ENC_PARTS": os.environ.get("MON_PIPE_ENC_PARTS", "default"),
                  "rays": a.rays, "hidden": a.hidden_layers, "iters_per_s": [round(r, 1) for r in rates], "loss": round(float(loss), 5)}))

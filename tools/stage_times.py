#!/usr/bin/env python
"""Per-stage times (mon_object_train_profiled: serial replay, CUDA event between kernels) and graph throughput at
several points of a training run: the gradient scatter and the optimizer sweep depend on how many samples still carry
gradient (early-stopped rays, background rays), so the cost of an iteration changes as the object converges.
usage: python tools/stage_times.py [--rays 4096] [--hidden-layers 1] [--frames 30]"""
import argparse, json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--hidden-layers", type=int, default=1)
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--at", default="0,50,200,500,1000,2000,5000")
a = ap.parse_args()
seq = syn.make_sequence(a.frames, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
g = core.NerfObject(ds, core.default_config(rays_per_batch=a.rays, n_hidden_layers=a.hidden_layers), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
done = 0
for upto in [int(x) for x in a.at.split(",")]:
    if upto > done:
        g.train(upto - done)
        done = upto
    st = g.train_profiled(20)
    done += 20
    g.prepare_train(20)
    g.train(20)                    # the 20 iterations right after: graph path, still in the same phase of training
    rate20 = 20.0 / (g.last_train_ms * 1e-3)
    done += 20
    loss = g.train(500)
    done += 500
    dout = g.last("dout") if False else None
    print(json.dumps({"after_iters": upto, "graph20_iters_per_s": round(rate20, 1), "graph_iters_per_s": round(500.0 / (g.last_train_ms * 1e-3), 1), "loss": round(float(loss), 5), "live_fraction": round(g.live_fraction, 4),
                      "stage_us": {k: round(v * 1e3, 2) for k, v in st.items()}}), flush=True)

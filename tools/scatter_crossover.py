#!/usr/bin/env python
"""Scatter time against the number of live samples for the two paths of k_scatter (MON_SCATTER_RESIDENT_MIN=0: always the
shared-memory resident path, -1: always global reductions): a fresh object is trained 5 profiled iterations at a time while the
early stop thins out the live samples; prints (live samples, scatter us) pairs per path and the crossover.
usage: python tools/scatter_crossover.py  (spawns itself once per path)"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1:
    sys.path.insert(0, str(ROOT))
    from ro_map_b200 import core, synthetic as syn
    seq = syn.make_sequence(30, 1)
    obj = seq.objects[0]
    ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
    for i in range(len(seq.rgb)):
        ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
    ds.sync()
    g = core.NerfObject(ds, core.default_config(rays_per_batch=4096), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
    g.set_bboxes(obj.boxes)
    rows = []
    for k in range(64):
        st = g.train_profiled(5 if k < 40 else 25)
        rows.append((int(round(g.live_fraction * 131072)), round(st["scatter"] * 1e3, 2)))
    print(json.dumps(rows))
    sys.exit(0)
curves = {}
for name, v in (("resident", "0"), ("global_reductions", "-1")):
    env = dict(os.environ, MON_SCATTER_RESIDENT_MIN=v)
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    curves[name] = json.loads(out)
    print(name, curves[name])
# crossover: first live count (descending) at which the reduction path is not slower than the resident one (same iteration index)
for (la, ta), (lb, tb) in zip(curves["resident"], curves["global_reductions"]):
    if tb <= ta:
        print(json.dumps({"crossover_live_samples_about": (la + lb) // 2, "resident_us": ta, "global_reductions_us": tb}))
        break

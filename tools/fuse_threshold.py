#!/usr/bin/env python
"""Which graph variant is faster at which live-sample count: a fresh object trained in calls of 25 iterations, per call the device
time per iteration, the kernels per iteration (6 = with the scatter kernel, 5 = scatter fused into the MLP kernel) and the live count
after the call — once per MON_SCATTER_FUSED_BELOW value (spawns itself).  usage: python tools/fuse_threshold.py"""
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
    g.prepare_train(25)
    rows = []
    for k in range(32):
        l0 = g.launch_count
        g.train(25)
        rows.append((int(round(g.live_fraction * 131072)), round(g.last_train_ms * 1e3 / 25, 1), (g.launch_count - l0) // 25))
    print(json.dumps(rows))
    sys.exit(0)
for thr in ("0", "16384", "32768", "49152", "65536", "140000"):
    env = dict(os.environ, MON_SCATTER_FUSED_BELOW=thr)
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    rows = json.loads(out)
    print(json.dumps({"fused_below": int(thr), "total_us_800_iters": round(sum(r[1] for r in rows) * 25), "calls (live after, us/iter, kernels/iter)": rows}))

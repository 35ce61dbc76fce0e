#!/usr/bin/env python
"""PSNR-vs-speed of the OPT-IN occupancy-grid mode against the default mode on the benchmark scene (800x800, R = 4096): both train
2000 iterations on the even keyframes, are rendered on the odd keyframes' object boxes (held out), and are timed over their last
500 iterations.  usage: python tools/occupancy_report.py [--res 64]"""
import argparse, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ro_map_b200 import core, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=64)
ap.add_argument("--alpha", type=float, default=0.01)
ap.add_argument("--seeds", type=int, default=3)
ap.add_argument("--modes", default="default,occupancy")
a = ap.parse_args()
seq = syn.make_sequence(30, 1, seed=1337)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
ds.add_frames(0, np.ascontiguousarray(np.stack(seq.rgb)), np.ascontiguousarray(np.stack(seq.instance)), np.ascontiguousarray(np.stack(seq.depth)), seq.poses)
ds.sync()
train_boxes = np.array([b for b in obj.boxes if int(b[0]) % 2 == 0])
test_boxes = [b for b in obj.boxes if int(b[0]) % 2 == 1]
def psnr_of(g):
    vals = []
    for b in test_boxes:
        fid, x, y, h, w = [int(v) for v in b]
        rgb, _, _ = g.render((fid, x, y, h, w), seq.poses[fid])
        gt = seq.rgb[fid][y:y + h, x:x + w].astype(np.float32) / 255.0
        inst = seq.instance[fid][y:y + h, x:x + w] == obj.instance_id
        if inst.sum() < 16:
            continue
        vals.append(-10.0 * np.log10(np.mean((rgb[inst] - gt[inst]) ** 2) + 1e-12))
    return float(np.mean(vals))
rows = []
for mode in a.modes.split(","):
    for seed in range(1337, 1337 + a.seeds):
        g = core.NerfObject(ds, core.default_config(rays_per_batch=4096), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id, seed=seed)
        g.set_bboxes(train_boxes)
        if mode == "occupancy":
            g.set_occupancy(a.res, warmup_iters=256, update_interval=16, alpha_threshold=a.alpha)
        g.train(1500)
        loss = g.train(500)
        row = {"mode": mode, "seed": seed, "iters_per_s_last_500": round(500.0 / (g.last_train_ms * 1e-3), 1), "loss": round(float(loss), 5), "psnr_held_out_db": round(psnr_of(g), 3)}
        if mode == "occupancy":
            row.update({k: round(v, 4) for k, v in g.occupancy_stats().items()})
        rows.append(row)
        print(json.dumps(row), flush=True)
        g.close()
for mode in a.modes.split(","):
    r = [x for x in rows if x["mode"] == mode]
    print(json.dumps({"summary": mode, "alpha_threshold": a.alpha if mode == "occupancy" else None, "grid": a.res if mode == "occupancy" else None, "iters_per_s": round(float(np.mean([x["iters_per_s_last_500"] for x in r])), 1), "psnr_db": round(float(np.mean([x["psnr_held_out_db"] for x in r])), 3)}))

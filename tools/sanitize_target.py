#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the training iteration (serial
injected path and graph path), a render and a density lattice on a tiny scene.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn

seq = syn.make_sequence(n_frames=4, n_objects=1, H=160, W=160, K=(222.222, 222.222, 80.0, 80.0))
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
for i in range(len(seq.poses)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
import os
# both scatter paths (MON_SCATTER_RESIDENT_MIN is read at object creation: 0 = always the shared-memory resident path, -1 = never),
# then the opt-in occupancy mode, and the block upload
blk = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
blk.add_frames(0, np.ascontiguousarray(np.stack(seq.rgb)), np.ascontiguousarray(np.stack(seq.instance)), np.ascontiguousarray(np.stack(seq.depth)), seq.poses)
blk.sync()
# raw 16-bit depth planes (converted by the batch kernel)
d16 = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
d16.set_depth_u16(1.0 / 5000.0)
d16.add_frames(0, np.ascontiguousarray(np.stack(seq.rgb)), np.ascontiguousarray(np.stack(seq.instance)),
               np.ascontiguousarray(np.stack([np.clip(np.rint(d * 5000.0), 0, 65535).astype(np.uint16) for d in seq.depth])), seq.poses)
d16.sync()
# (hidden layers, rays, MON_SCATTER_RESIDENT_MIN, occupancy grid, MON_SCATTER_FUSED): the last two rows run the graph variant
# without a scatter kernel (fused MLP kernel scatters; slim batch cluster beside the hash encode; two batch sets)
for nh, R, resident_min, occ, fused in ((1, 256, "0", 0, "-1"), (2, 128, "-1", 0, "-1"), (1, 256, "0", 16, "-1"), (1, 256, "-1", 0, "1"), (2, 128, "-1", 0, "1")):
    os.environ["MON_SCATTER_RESIDENT_MIN"] = resident_min
    os.environ["MON_SCATTER_FUSED"] = fused
    g = core.NerfObject(blk if occ else (d16 if fused == "1" else ds), core.default_config(rays_per_batch=R, n_hidden_layers=nh), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
    g.set_bboxes(obj.boxes)
    if occ:
        g.set_occupancy(occ, warmup_iters=4, update_interval=2, alpha_threshold=0.01)
    rng = np.random.default_rng(0)
    u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)
    print("injected", g.train_injected(u((R, 2)), u((R, 3)), u((R, 32))))
    print("graph", g.train(5), g.train(3), "live", g.live_fraction)
    fid, x, y, h, w = [int(v) for v in obj.boxes[0]]
    rgb, dep, mask = g.render((fid, x, y, min(h, 16), min(w, 16)), seq.poses[fid])
    print("render", float(mask.mean()), "lattice", float(g.density_grid((8, 8, 8)).mean()))
    sig = g.density_grid((12, 12, 12))
    m = g.extract_mesh(12, float(np.median(sig)))
    print("mesh", m["n_surface"], len(m["indices"]) // 3)
    if occ:
        print("occupancy", g.occupancy_stats())
    g.close()
rng = np.random.default_rng(3)
m = core.mesh_from_lattice(rng.uniform(0, 4, (11, 11, 11)).astype(np.float32), [-1, -1, -1], [1, 1, 1], 2.0)
print("lattice mesh", m["n_surface"], len(m["indices"]) // 3)
ds.close()
blk.close()
d16.close()
print("sanitize target done")

#!/usr/bin/env python
"""Range of the loss-scaled hash-grid gradient at the benchmarked shape: the shared-memory resident scatter accumulates in
32-bit fixed point with unit 2^-24 (kernels_scatter_smem.cu), i.e. +-128 per slice.  Prints max |gradient| per level at a
few points of a training run (injected-random iterations expose the gradient snapshot)."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn
seq = syn.make_sequence(30, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
R = 4096
cfg = core.default_config(rays_per_batch=R)
g = core.NerfObject(ds, cfg, obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
offs = core.grid_layout(cfg)[0] if hasattr(core, "grid_layout") else None
rng = np.random.default_rng(0)
u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)
done = 0
for at in (0, 5, 20, 100, 500):
    if at > done:
        g.train(at - done); done = at
    loss, n_in = g.train_injected(u((R, 2)), u((R, 3)), u((R, 32)))
    done += 1
    gr = g.state("grad")[g.n_mlp:]
    print(json.dumps({"after_iters": at, "loss": round(float(loss), 5), "live_fraction": round(g.live_fraction, 4), "max_abs_grid_grad_loss_scaled": float(np.abs(gr).max()),
                      "p999": float(np.quantile(np.abs(gr[gr != 0]), 0.999)) if (gr != 0).any() else 0.0, "nonzero": int((gr != 0).sum())}), flush=True)

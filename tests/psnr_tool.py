"""Paired PSNR runs: this repository's core vs the reference library (oracle/_ref), used by test_gpu_bench_config.py.
TEST INFRASTRUCTURE.  Held-out protocol of tools/psnr_compare.py: train on the even keyframes, render the odd keyframes' boxes with
the inference (EMA) weights of each side through ONE renderer (mon_object_render), PSNR = -10 log10(MSE) on the object's pixels."""
from __future__ import annotations

import numpy as np


def psnr_views(core, ds, cfg, seq, obj, weights, views):
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    g = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id)
    g.set_params(weights)
    vals = []
    for fid, x, y, h, w in views:
        rgb, _, _ = g.render((fid, x, y, h, w), seq.poses[fid], use_ema=False)
        gt = seq.rgb[fid][y:y + h, x:x + w].astype(np.float32) / 255.0
        m = seq.instance[fid][y:y + h, x:x + w] == obj.instance_id
        if m.sum() < 64:
            continue
        mse = float(((rgb.reshape(h, w, 3) - gt)[m] ** 2).mean())
        vals.append(-10.0 * np.log10(max(mse, 1e-12)))
    g.close()
    return float(np.mean(vals)), len(vals)


def paired_runs(core, ref_binding, syn, n_objects=2, n_seeds=10, iters=2000, size=800, frames=30, rays=4096, hidden=1, ref_repeat_seeds=0):
    f = 1111.11 * size / 800.0
    seq = syn.make_sequence(frames, n_objects, seed=1337, H=size, W=size, K=(f, f, size / 2.0, size / 2.0))
    ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
    for i in range(len(seq.rgb)):
        ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
    ds.sync()
    cfg = core.default_config(rays_per_batch=rays, n_hidden_layers=hidden)
    lib = ref_binding.RefLib()
    rows = []

    def ref_run(obj, train, seed):
        r = ref_binding.RefModel(hidden, seed, lib)
        r.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, train, obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id, True, rays)
        _, _, loss, _ = r.train(iters)
        ema = r.get(2)
        r.close()
        return ema, loss

    for obj in seq.objects:
        train = [b for b in obj.boxes if b[0] % 2 == 0]
        held = [b for b in obj.boxes if b[0] % 2 == 1]
        for k in range(n_seeds):
            seed = 1337 + k
            g = core.NerfObject(ds, cfg, obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id, seed)
            g.set_bboxes(train)
            loss_o = g.train(iters)
            ours = g.state("ema")
            g.close()
            ref, loss_r = ref_run(obj, train, seed)
            p_o, n_views = psnr_views(core, ds, cfg, seq, obj, ours, held)
            p_r, _ = psnr_views(core, ds, cfg, seq, obj, ref, held)
            row = {"object": int(obj.instance_id), "kind": obj.kind, "seed": seed, "iters": iters, "held_out_views": n_views, "image": f"{size}x{size}",
                   "psnr_ours_db": round(p_o, 3), "psnr_reference_db": round(p_r, 3), "delta_db": round(p_o - p_r, 3),
                   "loss_ours": round(float(loss_o), 6), "loss_reference": round(float(loss_r), 6)}
            if k < ref_repeat_seeds:   # the reference against ITSELF: same seed, same inputs, second run
                ref2, _ = ref_run(obj, train, seed)
                p_r2, _ = psnr_views(core, ds, cfg, seq, obj, ref2, held)
                row["psnr_reference_rerun_db"] = round(p_r2, 3)
                row["ref_rerun_delta_db"] = round(p_r2 - p_r, 3)
            rows.append(row)
    ds.close()
    return rows

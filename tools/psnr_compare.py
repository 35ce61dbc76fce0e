#!/usr/bin/env python
"""PSNR of held-out views: this repository's core vs the reference Core (SURVEY.md §8d "PSNR vs reference (±0.1 dB)").

Both sides train the same objects of the synthetic 'room' stand-in for the same number of iterations on the EVEN
keyframes (each with its own random stream: ours the counter-based hash, the reference cuRAND XORWOW), then the ODD
keyframes' boxes are rendered with the inference (EMA) weights of each side through one and the same renderer
(mon_object_render, so that only the trained weights differ) and compared with the ground-truth pixels of the object
(instance mask) — PSNR = -10 log10(MSE) on float RGB in [0, 1] (TCNN/scripts/common.py:32), mean over views.

    python tools/psnr_compare.py [--iters 2000] [--objects 2] [--seeds 3] [--rays 4096] [--frames 30] [--size 400]

TEST / MEASUREMENT TOOL: uses oracle/ref (the reference's vendored tiny-cuda-nn built under oracle/_ref) as the checker.
Prints one JSON line per (object, seed) and a summary line.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "ref"))


def psnr_views(core, ds, cfg, seq, obj, weights_fp16_as_f32, views):
    """Render `views` with the given fp16 weight vector (installed as the training weights of a scratch object) and
    return the mean PSNR over the object's pixels."""
    bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
    g = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id)
    g.set_params(weights_fp16_as_f32)
    vals = []
    for fid, x, y, h, w in views:
        rgb, dep, mask = g.render((fid, x, y, h, w), seq.poses[fid], use_ema=False)
        gt = seq.rgb[fid][y:y + h, x:x + w].astype(np.float32) / 255.0
        m = seq.instance[fid][y:y + h, x:x + w] == obj.instance_id
        if m.sum() < 64:
            continue
        mse = float(((rgb.reshape(h, w, 3) - gt)[m] ** 2).mean())
        vals.append(-10.0 * np.log10(max(mse, 1e-12)))
    g.close()
    return float(np.mean(vals)), len(vals)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--objects", type=int, default=2)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--size", type=int, default=400)
    ap.add_argument("--hidden-layers", type=int, default=1)
    ap.add_argument("--inject", action="store_true", help="train our side with host-generated (numpy PCG64) random arrays through the "
                    "parity hook instead of the in-kernel counter-based generator (isolates the effect of the random stream)")
    args = ap.parse_args()

    from ro_map_b200 import core, synthetic as syn
    import ref_binding

    s = args.size
    f = 1111.11 * s / 800.0
    seq = syn.make_sequence(args.frames, args.objects, seed=1337, H=s, W=s, K=(f, f, s / 2.0, s / 2.0))
    ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
    for i in range(len(seq.rgb)):
        ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
    ds.sync()
    cfg = core.default_config(rays_per_batch=args.rays, n_hidden_layers=args.hidden_layers)
    lib = ref_binding.RefLib()
    rows = []
    for obj in seq.objects:
        train = [b for b in obj.boxes if b[0] % 2 == 0]
        held = [b for b in obj.boxes if b[0] % 2 == 1]
        bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
        for k in range(args.seeds):
            seed = 1337 + k
            g = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id, seed)
            g.set_bboxes(train)
            if args.inject:
                rng = np.random.default_rng(seed)
                R = args.rays
                for _ in range(args.iters):
                    # curandGenerateUniform's interval is (0, 1]
                    xy, col, dt = (1.0 - rng.random(n, dtype=np.float32) for n in (2 * R, 3 * R, 32 * R))
                    loss_o, _ = g.train_injected(xy, col, dt)
            else:
                loss_o = g.train(args.iters)
            ours_ema = g.state("ema")
            g.close()
            r = ref_binding.RefModel(args.hidden_layers, seed, lib)
            r.scene(seq.rgb, seq.instance, seq.depth, seq.poses, seq.H, seq.W, seq.K, train, obj.Tow, bmin, bmax, obj.instance_id, True, args.rays)
            _, _, loss_r, _ = r.train(args.iters)
            ref_kernels = "RO-MAP's own nerf_model.cu kernels + unmodified tiny-cuda-nn" if r.is_genuine() else "glue kernels restated + unmodified tiny-cuda-nn"
            ref_ema = r.get(2)
            r.close()
            p_o, n_views = psnr_views(core, ds, cfg, seq, obj, ours_ema, held)
            p_r, _ = psnr_views(core, ds, cfg, seq, obj, ref_ema, held)
            row = {"object": obj.instance_id, "kind": obj.kind, "seed": seed, "iters": args.iters, "held_out_views": n_views,
                   "psnr_ours_db": round(p_o, 3), "psnr_reference_db": round(p_r, 3), "delta_db": round(p_o - p_r, 3),
                   "loss_ours": round(float(loss_o), 6), "loss_reference": round(float(loss_r), 6)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    d = np.array([r["delta_db"] for r in rows])
    print(json.dumps({"summary": True, "runs": len(rows), "mean_psnr_ours_db": round(float(np.mean([r["psnr_ours_db"] for r in rows])), 3),
                      "mean_psnr_reference_db": round(float(np.mean([r["psnr_reference_db"] for r in rows])), 3),
                      "mean_delta_db": round(float(d.mean()), 3), "std_delta_db": round(float(d.std()), 3),
                      "ours_random_stream": "numpy (injected)" if args.inject else "in-kernel", "rays_per_batch": args.rays, "n_hidden_layers": args.hidden_layers, "image": f"{s}x{s}", "keyframes": args.frames,
                      "reference": ref_kernels}))


if __name__ == "__main__":
    main()

"""Diagnostic: after ONE injected iteration, where do the CUDA weights differ from the oracle's (bit level)?"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tests')
from oracle import mon_oracle as orc
from ro_map_b200 import core, synthetic as syn
seq = syn.make_sequence(n_frames=4, n_objects=1, H=160, W=160, K=(222.222, 222.222, 80.0, 80.0))
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
for i in range(len(seq.poses)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
R = 256
cfg = core.default_config(rays_per_batch=R)
bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
g = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id)
g.set_bboxes(obj.boxes)
o = orc.OracleObject(orc.default_config(), R, 32, obj.Tow, bmin, bmax, obj.instance_id, True, n_threads=4)
frames = orc.Frames(seq.rgb, seq.instance, seq.depth, seq.poses)
rng = np.random.default_rng(0)
u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)
for it in range(2):
    sxy, col, dt = u((R, 2)), u((R, 3)), u((R, 32))
    lg, ng = g.train_injected(sxy, col, dt)
    lo, no = o.train_iter(obj.boxes, frames, seq.H, seq.W, seq.K, sxy, col, dt)
    print("iter", it, "loss", lg, lo, "rays equal", np.array_equal(g.last("rays"), o.last("rays")), "enc equal", np.array_equal(g.last("enc"), o.last("enc")),
          "enc mismatches", int((g.last("enc") != o.last("enc")).sum()))
    n_mlp = g.n_mlp
    for name in ("grad", "master", "params", "adam_m", "adam_v", "param_steps", "ema"):
        a, b = g.state(name), o.state(name)
        d = a != b
        print(f"   {name:12s} mismatches mlp {int(d[:n_mlp].sum()):6d} grid {int(d[n_mlp:].sum()):8d}  max abs diff mlp {np.abs(a[:n_mlp]-b[:n_mlp]).max():.3e} grid {np.abs(a[n_mlp:]-b[n_mlp:]).max():.3e}")
    gm, om = g.state("master"), o.state("master")
    gs, os_ = g.state("param_steps"), o.state("param_steps")
    same_touch = (gs == os_)
    d = (gm != om) & same_touch
    idx = np.nonzero(d[n_mlp:])[0][:6]
    gg, og = g.state("grad"), o.state("grad")
    for i in idx:
        j = n_mlp + i
        print("      grid param", i, "master", gm[j], om[j], "grad", gg[j], og[j], "m", g.state("adam_m")[j], o.state("adam_m")[j])

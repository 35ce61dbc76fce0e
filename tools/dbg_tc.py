"""Cross-checks the tcgen05 MLP kernel against the mma.sync validation kernel on identical state (GPU)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
from ro_map_b200 import core, synthetic as syn

seq = syn.make_sequence(n_frames=6, n_objects=2, seed=1337, H=200, W=200, K=(277.7775, 277.7775, 100.0, 100.0))
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
for i in range(len(seq.poses)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = core.default_config(rays_per_batch=R)
bmin, bmax = -1.1 * obj.half, 1.1 * obj.half
rng = np.random.default_rng(0)
u = lambda shape: (1.0 - rng.random(shape, dtype=np.float32)).astype(np.float32)
a = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id); a.set_mlp_impl(1); a.set_bboxes(obj.boxes)
a.train(30)
start = a.state("master")
res = {}
sxy, col, dt = u((R, 2)), u((R, 3)), u((R, 32))
for impl in (1, 0):
    g = core.NerfObject(ds, cfg, obj.Tow, bmin, bmax, obj.instance_id); g.set_mlp_impl(impl); g.set_bboxes(obj.boxes)
    g.set_params(start)
    loss, n_in = g.train_injected(sxy, col, dt)
    res[impl] = dict(loss=loss, n_in=n_in, out=g.last("out"), rgb=g.last("rgb_rays"), dout=g.last("dout"), d_enc=g.last("d_enc"),
                     grad=g.state("grad")[:g.n_mlp], ms=g.last_train_ms)
    print("impl", impl, "loss", loss, "n_in", n_in, "ms", g.last_train_ms)
for k in ("out", "rgb", "dout", "d_enc", "grad"):
    x, y = res[1][k], res[0][k]
    print(f"{k:6s} max|wmma|={np.abs(x).max():.4e} max|tc|={np.abs(y).max():.4e} max|diff|={np.abs(x-y).max():.4e} "
          f"rel={np.abs(x-y).max()/max(np.abs(x).max(),1e-30):.3e} nan={np.isnan(y).sum()}")
gw = res[1]["grad"]; gt = res[0]["grad"]
print("grad W_in corr", np.corrcoef(gw[:2048], gt[:2048])[0, 1], "W_out corr", np.corrcoef(gw[2048:], gt[2048:])[0, 1])
print("W_out rows 0-3 wmma", gw[2048:2048+4], "tc", gt[2048:2048+4])

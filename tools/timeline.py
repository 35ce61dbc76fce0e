#!/usr/bin/env python
"""Timeline of the iteration graph on the device: which kernels of the forked iteration graph really overlap.

Needs the instrumented build (every CTA folds %globaltimer into a per-(iteration, kernel) [first start, last end] slot,
ro_map_b200/csrc/mon_timeline.cuh), which `--build` cross-compiles into ro_map_b200/_build_tl/ without touching the
product library:

    MON_EXTRA_NVCC_FLAGS=-DMON_TIMELINE python tools/timeline.py --build      # here (no GPU needed)
    python tools/timeline.py [--at 5,400]                                       # on the B200 box

Prints, for a few iterations after each --at count (5 = the dense start-up phase of a fresh object, 400 = steady state), start/end of every kernel in microseconds relative to the start of that
iteration's fused MLP kernel, plus the iteration period."""
import argparse, ctypes as C, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
TL_DIR = ROOT / "ro_map_b200" / "_build_tl"
TL_LIB = TL_DIR / "libmon_b200_tl.so"

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--hidden-layers", type=int, default=1)
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--at", default="5,400", help="iteration counts after which a 32-iteration graph is traced")
a = ap.parse_args()

if a.build:
    assert "-DMON_TIMELINE" in os.environ.get("MON_EXTRA_NVCC_FLAGS", ""), "set MON_EXTRA_NVCC_FLAGS=-DMON_TIMELINE"
    from ro_map_b200 import build
    build.BUILD, build.LIB = TL_DIR, TL_LIB
    TL_DIR.mkdir(exist_ok=True)
    print(build.build(force=True))
    sys.exit(0)

from ro_map_b200 import _capi
_capi.LIB_PATH = TL_LIB
from ro_map_b200 import core, synthetic as syn
lib = _capi.load()
ITERS, KINDS = 64, 16
# S = global-reduction scatter kernel [first CTA start, last CTA end], S1 = shared-memory resident scatter kernel (only the one that took the iteration leaves a mark)
NAMES = ["B", "P", "E0", "E1", "E2", "E3", "M", "S", "S1", "S2", "S3", "O0", "O1", "O2", "O3", "O"]
TABLES = ["batch", "encode", "mlp", "optim", "scatter_smem"]


def reset():
    for t in TABLES:
        assert getattr(lib, f"mon_debug_tl_{t}_reset")() == 0


def read():
    start = np.full((ITERS, KINDS), np.iinfo(np.uint64).max, np.uint64)
    end = np.zeros((ITERS, KINDS), np.uint64)
    for t in TABLES:
        buf = np.zeros((ITERS, KINDS, 2), np.uint64)
        assert getattr(lib, f"mon_debug_tl_{t}_read")(buf.ctypes.data_as(C.c_void_p)) == 0
        start = np.minimum(start, buf[..., 0])
        end = np.maximum(end, buf[..., 1])
    return start, end


seq = syn.make_sequence(a.frames, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
g = core.NerfObject(ds, core.default_config(rays_per_batch=a.rays, n_hidden_layers=a.hidden_layers), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
done = 0
for at in [int(x) for x in a.at.split(",")]:
    if at > done:
        g.train(at - done)
        done = at
    reset()
    g.train(32)                    # one 32-iteration graph: iterations at .. at+31
    ms = g.last_train_ms
    start, end = read()
    print(json.dumps({"after_iters": at, "graph_us_per_iter": round(ms * 1e3 / 32, 2)}))
    periods = []
    for it in range(at + 12, at + 18):
        s = it % ITERS
        m0 = int(start[s, 6])
        row = []
        for k in sorted(range(KINDS), key=lambda k: int(start[s, k]) if end[s, k] else 1 << 62):
            if end[s, k] == 0:
                continue
            row.append(f"{NAMES[k]}[{(int(start[s, k]) - m0) / 1e3:.1f},{(int(end[s, k]) - m0) / 1e3:.1f}]")
        nxt = int(start[(it + 1) % ITERS, 6])
        periods.append((nxt - m0) / 1e3)
        print(f"iter {it}: period {(nxt - m0) / 1e3:.1f} us  " + " ".join(row))
    print(json.dumps({"after_iters": at, "mean_period_us": round(float(np.mean(periods)), 2)}))
    # every iteration of the traced graph: period (M start to next M start) and kernel durations, to see drift inside the window
    rows = []
    for it in range(at, at + 31):
        s0, s1 = it % ITERS, (it + 1) % ITERS
        if not end[s0, 6] or not end[s1, 6]:
            continue
        dur = lambda k: round((int(end[s0, k]) - int(start[s0, k])) / 1e3, 1) if end[s0, k] else None   # noqa: E731
        rows.append({"it": it, "period": round((int(start[s1, 6]) - int(start[s0, 6])) / 1e3, 1), "E": dur(2), "M": dur(6), "S": dur(7), "S1": dur(8), "O": dur(15)})
    print("per-iteration: " + json.dumps(rows))
    done += 32
    # per-CTA phases of the shared-memory resident scatter for the traced iteration with (iter % 64) == 20
    try:
        buf = np.zeros((256, 3), np.uint64)
        assert lib.mon_debug_tl_sr_cta_read(buf.ctypes.data_as(C.c_void_p)) == 0
        live = buf[:, 2] > 0
        if live.any():
            t0 = int(buf[live, 0].min())
            rel = (buf[live].astype(np.int64) - t0) / 1e3
            print("resident scatter per CTA: start / first job accumulated / done (us after the first CTA start)")
            for k in range(3):
                print("  " + " ".join(f"{v:.1f}" for v in rel[:, k]))
    except AttributeError:
        pass

# per-CTA load balance of the encode kernel (iteration 404 + 16 = slot 20 of the instrumented replay)
try:
    buf = np.zeros((256, 2), np.uint64)
    assert lib.mon_debug_tl_enc_cta_read(buf.ctypes.data_as(C.c_void_p)) == 0
    live = buf[:, 1] > 0
    t0 = int(buf[live, 0].min())
    dur = (buf[live, 1].astype(np.int64) - t0) / 1e3
    load = (buf[live, 0].astype(np.int64) - t0) / 1e3
    print("encode per-CTA done time after the first table became resident (us), CTA 0..%d:" % (live.sum() - 1))
    print(" ".join(f"{d:.1f}" for d in dur))
    print("table resident (us): " + " ".join(f"{d:.1f}" for d in load))
except AttributeError:
    pass

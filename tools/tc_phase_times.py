"""Phase timing of the fused MLP kernel (CTA 0): build with MON_EXTRA_NVCC_FLAGS=-DMON_TC_STAMPS, then run this on a GPU."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from ro_map_b200 import _capi
from pathlib import Path
_capi.LIB_PATH = Path('/root/repo/ro_map_b200/_build_tl2/libmon_b200_stamps.so')
from ro_map_b200 import core, synthetic as syn
seq = syn.make_sequence(n_frames=12, n_objects=1, seed=1337, H=400, W=400, K=(555.555, 555.555, 200.0, 200.0))
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.poses), True)
for i in range(len(seq.poses)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
g = core.NerfObject(ds, core.default_config(), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
g.train(600)
print("stage ms", g.train_profiled(20))
lib = _capi.load()
st = (C.c_ulonglong * 64)()
assert lib.mon_debug_tc_stamps(st) == 0
s = np.array(st[:], dtype=np.int64)
names = {0: "entry", 1: "setup done", 2: "weights loaded", 40: "tiles done", 41: "partials written", 42: "exit"}
tile = ["tile start", "enc stored+sync", "MMA1 issued", "ray inputs loaded", "MMA1 done", "hidden epilogue", "MMA2 issued", "MMA2 done",
        "render+loss epilogue", "MMA3 issued", "MMA3 done", "dhid epilogue", "MMA4+wgrad issued", "MMA4+wgrad done", "d_enc stored"]
t0 = s[0]
for i in (0, 1, 2):
    print(f"{names[i]:28s} {s[i]-t0:8d} cyc")
for base in (4, 20):
    prev = None
    for k, n in enumerate(tile):
        v = s[base + k]
        if v == 0: continue
        print(f"  [{'tile A' if base == 4 else 'tile B'}] {n:24s} {v-t0:8d} cyc" + (f"  (+{v-prev})" if prev else ""))
        prev = v
for i in (40, 41, 42):
    print(f"{names[i]:28s} {s[i]-t0:8d} cyc")
ct = (C.c_ulonglong * (1024 * 3))()
assert lib.mon_debug_tc_ctas(ct) == 0
a = np.array(ct[:], dtype=np.int64).reshape(1024, 3)[:g.R // 4 if g.R // 4 < 592 else 592]
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
print(f"CTAs {len(a)}: first start 0, last start {(a[:,0].max()-t0)/1e3:.2f} us, first end {(a[:,1].min()-t0)/1e3:.2f} us, last end {(a[:,1].max()-t0)/1e3:.2f} us")
import collections
per_sm = collections.defaultdict(list)
for s0, e0, sm in a: per_sm[int(sm)].append((s0 - t0, e0 - t0))
conc = []
for sm, iv in per_sm.items():
    starts = sorted(x[0] for x in iv)
    # CTAs of this SM that started within 1 us of the kernel start = resident in the first wave
    conc.append(sum(1 for x in starts if x < 1500))
print("SMs used", len(per_sm), "CTAs/SM in first wave: min", min(conc), "max", max(conc), "hist", collections.Counter(conc))
print("CTA duration us: mean %.2f max %.2f" % (((a[:,1]-a[:,0]).mean())/1e3, ((a[:,1]-a[:,0]).max())/1e3))

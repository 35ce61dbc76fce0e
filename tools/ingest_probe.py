"""Keyframe ingest latency of mon_dataset_add_frame (800x800, u8 RGB + u8 instance + f32 depth) from pageable and from
pinned host buffers, with the GPU idle and while an object is training on another stream."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from ro_map_b200 import core, synthetic as syn
seq = syn.make_sequence(30, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, 64, True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
g = core.NerfObject(ds, core.default_config(), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
g.set_bboxes(obj.boxes)
g.train(100)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
prgb, pinst, pdep = pin(seq.rgb[0]), pin(seq.instance[0]), pin(seq.depth[0])
def probe(label, rgb, inst, dep, busy):
    ts = []
    for k in range(12):
        if busy:
            g.train_async(2000)          # ~125 ms of queued graph replays
        t0 = time.perf_counter()
        ds.add_frame(32 + k % 16, rgb, inst, dep, seq.poses[0])
        ts.append((time.perf_counter() - t0) * 1e3)
        if busy:
            g.sync()
    ds.sync()
    print(f"{label:34s} median {np.median(ts[2:]):7.3f} ms  max {max(ts[2:]):7.3f} ms")
probe("pageable, GPU idle", seq.rgb[0], seq.instance[0], seq.depth[0], False)
probe("pinned,   GPU idle", prgb, pinst, pdep, False)
probe("pageable, object training", seq.rgb[0], seq.instance[0], seq.depth[0], True)
probe("pinned,   object training", prgb, pinst, pdep, True)

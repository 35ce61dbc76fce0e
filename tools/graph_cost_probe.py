#!/usr/bin/env python
"""Host cost of object creation and of instantiating the iteration graphs (the calls that hold the CUDA driver while other
threads ingest keyframes in online mode): wall time of mon_object_create and of mon_object_prepare_train for several lengths."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ro_map_b200 import core, synthetic as syn
seq = syn.make_sequence(8, 1)
obj = seq.objects[0]
ds = core.Dataset(0, *seq.K, seq.H, seq.W, len(seq.rgb), True)
for i in range(len(seq.rgb)):
    ds.add_frame(i, seq.rgb[i], seq.instance[i], seq.depth[i], seq.poses[i])
ds.sync()
out = {}
for trial in range(2):
    t0 = time.perf_counter()
    g = core.NerfObject(ds, core.default_config(), obj.Tow, -1.1 * obj.half, 1.1 * obj.half, obj.instance_id)
    out[f"create_ms_{trial}"] = round((time.perf_counter() - t0) * 1e3, 2)
    g.set_bboxes(obj.boxes)
    for n in (1, 20, 50, 64, 52):
        t0 = time.perf_counter()
        g.prepare_train(n)
        out[f"prepare_{n}_ms_{trial}"] = round((time.perf_counter() - t0) * 1e3, 2)
    g.close()
print(json.dumps(out))

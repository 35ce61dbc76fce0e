#!/usr/bin/env python
"""Host-to-device throughput of the keyframe upload pattern (pinned memory): 30 frames x (1.92 MB rgb + 0.64 MB instance + 2.56 MB
depth) as 90 copies on one stream (what mon_dataset_add_frame issues), as 30 packed copies, as one copy, and split over two streams."""
import json, time
import torch
px = 800 * 800
sizes = [px * 3, px, px * 4]
n = 30
total = n * sum(sizes)
host = torch.empty(total, dtype=torch.uint8).pin_memory()
dev = torch.empty(total, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(plan):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for st, off, nb in plan:
            with torch.cuda.stream(st):
                dev[off:off + nb].copy_(host[off:off + nb], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return round(best * 1e3, 3), round(total / best / 1e9, 1)
plans = {}
off, p90, p90_2 = 0, [], []
for f in range(n):
    for k, nb in enumerate(sizes):
        p90.append((s1, off, nb)); p90_2.append((s1 if k < 2 else s2, off, nb)); off += nb
plans["90 copies, 1 stream"] = p90
plans["90 copies, depth on a 2nd stream"] = p90_2
per = sum(sizes)
plans["30 packed copies, 1 stream"] = [(s1, f * per, per) for f in range(n)]
plans["30 packed copies, alternating 2 streams"] = [((s1, s2)[f & 1], f * per, per) for f in range(n)]
plans["1 copy"] = [(s1, 0, total)]
plans["2 halves, 2 streams"] = [(s1, 0, total // 2), (s2, total // 2, total - total // 2)]
print(json.dumps({k: dict(zip(("ms", "GB/s"), run(v))) for k, v in plans.items()}, indent=1))

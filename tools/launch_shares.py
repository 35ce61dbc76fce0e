#!/usr/bin/env python
"""Per-kernel shares of the timed window from an ncu launch list (--metrics gpu__time_duration.sum --csv) of the driver-shaped
bench command: the launches of iterations 5..25 of the first object (the first 5 x 6 + prologue launches are the warm-up call).
usage: python tools/launch_shares.py profiles/r9_launches_20_5.csv > profiles/r9_launch_shares_20_5.txt"""
import csv, re, sys
from collections import OrderedDict
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else v * (1e3 if unit in ("msecond", "ms") else 1.0)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), us))
train = [(k, t) for k, t in rows if k.startswith(("k_generate_batch", "k_sample_points", "k_encode_forward", "k_mlp_train_tc", "k_scatter", "k_optimizer_sweep"))]
# warm-up call = 5 iterations (30 launches), timed call = 20 iterations (120 launches)
window = train[30:150]
agg = OrderedDict()
for k, t in window:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list of `python bench.py --steps 20 --warmup 5` (gpu__time_duration.sum, --clock-control none, graph nodes profiled one by one: cold caches,")
print(f"# serialised) — launches of iterations 5..25 of the first object = the timed window; per-kernel mean duration and share of the step")
print(f"# source: {path} ({len(rows)} launches captured)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} launches {n:4d}  mean {t / n:8.2f} us  share {100 * t / tot:5.1f} %")
print(f"{'sum per iteration':28s} {tot / 20:8.2f} us (serialised, cold caches; the graph's device time per iteration is in the bench line)")

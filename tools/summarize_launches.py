#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0.0, 0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    agg[name][0] += v
    agg[name][1] += 1
    total += v
print(f"total kernel time {total / 1e6:.3f} ms over {sum(v[1] for v in agg.values())} launches")
print(f"{'share':>7s} {'ms':>9s} {'launches':>8s} {'avg us':>9s}  kernel")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / total * 100:6.2f}% {t / 1e6:9.3f} {n:8d} {t / n / 1e3:9.1f}  {name[:110]}")

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: summarize_launches.py launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"]
    for pre in ("void ", "at::native::"):
        name = name.replace(pre, "")
    name = name[:90]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {len(rows)} launches, {tot:.1f} us total (per-launch times are cold-cache, serialised)")
print(f"{'us':>10} {'n':>5} {'share':>7}  kernel")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:10.1f} {a[0]:5d} {100 * a[1] / tot:6.1f}%  {n}")

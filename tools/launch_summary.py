#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n] > profiles/xxx.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, ns, grid, block in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void\s+", "", short)
    d = agg.setdefault((short, grid, block), [0, 0.0])
    d[0] += 1; d[1] += ns
tot = sum(v[1] for v in agg.values())
print("| kernel | grid | block | launches | total ms | mean us | share |")
print("|---|---|---|---:|---:|---:|---:|")
for (name, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %s | %s | %d | %.3f | %.1f | %.1f%% |" % (name, grid, block, n, ns / 1e6, ns / n / 1e3, 100 * ns / tot))
print("\ntotal %.3f ms over %d launches" % (tot / 1e6, len(rows)))

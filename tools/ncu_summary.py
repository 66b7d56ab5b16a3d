#!/usr/bin/env python
"""Markdown table of the metrics that matter from an `ncu --set full` capture, one row per kernel (mean over
its profiled launches).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv > profiles/xxx.md"""
import csv, re, sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}
M = [("ms", "gpu__time_duration.sum", 1.0),
     ("DRAM rd GB", "dram__bytes_read.sum", 1.0), ("DRAM wr GB", "dram__bytes_write.sum", 1.0),
     ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0), ("L2 thr %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1.0), ("L1 thr %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
     ("IPC", "sm__inst_executed.avg.per_cycle_elapsed", 1.0), ("warp-inst G", "smsp__inst_executed.sum", 1e-9),
     ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0), ("regs", "launch__registers_per_thread", 1.0),
     ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
     ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1.0)]


def unit_scale(u, want_giga):
    u = u.lower()
    if not want_giga:
        return 1.0
    return {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0, "tbyte": 1e3}.get(u, 1.0)


def tscale(u):
    return {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u.lower(), 1.0)


agg = OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    key = (name, r[col["Grid Size"]])
    vals = []
    for label, k, sc in M:
        if k not in col or r[col[k]] == "":
            vals.append(None); continue
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]]
        if label == "ms": v *= tscale(u)
        elif "GB" in label: v *= unit_scale(u, True)
        else: v *= sc
        vals.append(v)
    agg.setdefault(key, []).append(vals)
print("| kernel | grid | n | " + " | ".join(m[0] for m in M) + " |")
print("|---|---|---:|" + "---:|" * len(M))
for (name, grid), lst in agg.items():
    means = []
    for i in range(len(M)):
        xs = [v[i] for v in lst if v[i] is not None]
        means.append(sum(xs) / len(xs) if xs else None)
    print("| `%s` | %s | %d | " % (name, grid, len(lst)) + " | ".join("-" if m is None else ("%.3f" % m if abs(m) < 100 else "%.1f" % m) for m in means) + " |")

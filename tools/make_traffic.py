#!/usr/bin/env python
"""
profiles/traffic.json from an `ncu --set full` raw CSV of ONE apply of the fused recipe: DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) per kernel, keyed by the labels bench.py's per-kernel timing uses.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/make_traffic.py raw.csv cfg3 16 "profiles/r02_sN_cfg3.md (ncu --set full, one apply)" [more: raw.csv workload coils source ...]

The launches of one apply arrive in order: expand, pass (y fwd), pass (z fwd), kb_gather, runs (+seg, +fold),
pass (z inv), pass (y inv), combine; the four strided passes are told apart by that order.
"""
import csv
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scale(u):
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u.lower(), 1.0)


def one(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}
    out, npass = {}, 0
    order = ["fft_pass[y fwd]", "fft_pass[z fwd]", "fft_pass[z inv]", "fft_pass[y inv]"]
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[k]].replace(",", "")) * scale(units[col[k]])
        if "sense_expand" in name:
            if "sense_expand_pk[x]" in out:          # a second apply starts: stop
                break
            label = "sense_expand_pk[x]"
        elif "sense_combine" in name:
            label = "sense_combine_pk[x]"
        elif "kb_gather" in name:
            label = "kb_gather"
        elif "csrmm_runs" in name:
            label = "csrmm_runs"
        elif "kb_blocks" in name:
            label = "kb_blocks"
        elif "fft_pk" in name or "fft_il_pass" in name or "fft_spec" in name:
            label = order[min(npass, 3)]; npass += 1
        else:
            continue
        out[label] = out.get(label, 0.0) + b
    return {k: int(v) for k, v in out.items()}


def main():
    args = sys.argv[1:]
    path = os.path.join(REPO, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table = {k: v for k, v in table.items() if isinstance(v, dict) and "kernels" in v}
    for i in range(0, len(args), 4):
        raw, workload, coils, source = args[i:i + 4]
        table["%s:coils%d" % (workload, int(coils))] = {"kernels": one(raw), "source": source}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1)[:1500])


if __name__ == "__main__":
    main()

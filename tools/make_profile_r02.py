#!/usr/bin/env python
"""profiles/<name>.md from the files a tools/gpu_r2_s12.sh-style session left in gpurun_out/ (tests, smoke, bench lines,
launch list, ncu --set full raw CSVs at 16 and at 2 coils).
usage: python tools/make_profile_r02.py <tag> <name> "<title>" """
import json, os, subprocess, sys

tag, name, title = sys.argv[1], sys.argv[2], sys.argv[3]
G = "gpurun_out"


def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def last_json(path):
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


out = ["# %s" % title, ""]
tests = open(os.path.join(G, tag + "_tests.log")).read().strip().splitlines()
passed = [l for l in tests if " passed" in l or " failed" in l]
out += ["GPU tests of this build (`pytest tests -m gpu`): `%s`" % (passed[-1].strip("= ") if passed else "?"), ""]
p = os.path.join(G, tag + "_smoke.log")
if os.path.exists(p):
    out += ["`__graft_entry__.smoke()`:", "```"] + [l for l in open(p).read().splitlines() if l.startswith("smoke")] + ["```", ""]
for label, f in (("fused recipe (default), `python bench.py --check --check-tree`", "_bench_cfg3.json"),
                 ("per-GPU shard of a 2-GPU run on one GPU, `--coils 8` (development flag)", "_bench_coils8.json"),
                 ("per-GPU shard of a 4-GPU run, `--coils 4`", "_bench_coils4.json"),
                 ("per-GPU shard of an 8-GPU run, `--coils 2`", "_bench_coils2.json")):
    p = os.path.join(G, tag + f)
    if os.path.exists(p) and os.path.getsize(p):
        d = last_json(p)
        out += ["### " + label, "",
                "%.2f applies/s device-resident (%.3f ms per apply), %.2f applies/s end to end (%s)" %
                (d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d["e2e"].get("paths_timed"))), ""]
        out += ["| kernel | ms | launches | compulsory GB | frac (compulsory) | ncu DRAM GB | frac (ncu DRAM) | frac (replaced call) |",
                "|---|---:|---:|---:|---:|---:|---:|---:|"]
        for k in d["kernels"]:
            out.append("| %s | %.3f | %d | %.2f | %.3f | %s | %s | %.3f |" % (
                k["kernel"], k["ms"], k["launches"], k["bytes"] / 1e9, k["frac"],
                "%.2f" % (k["dram_bytes"] / 1e9) if k.get("dram_bytes") else "-",
                "%.3f" % k["frac_dram"] if k.get("frac_dram") else "-", k["frac_replaced_call"]))
        out += ["", "check: `%s`" % json.dumps(d.get("check")), "", "setup: `%s`, clocks `%s`" % (json.dumps(d.get("setup")), json.dumps(d.get("clocks"))), ""]
        if d.get("cpu_baseline"):
            cb = {k: v for k, v in d["cpu_baseline"].items() if k in ("value", "unit", "cores", "kind", "sample", "seconds_per_apply", "extrapolated")}
            out += ["cpu_baseline: `%s`" % json.dumps(cb), ""]
p = os.path.join(G, tag + "_launches_cfg3.csv")
if os.path.exists(p):
    out += ["## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
            "(cold-cache, serialised: compare shares; includes the one-time device construction of the operator)", "",
            sh("python tools/launch_summary.py %s" % p).strip(), ""]
for label, f in (("16 coils (one GPU)", "_raw_cfg3.csv"), ("2 coils (the per-GPU shard of an 8-GPU run)", "_raw_coils2.csv")):
    p = os.path.join(G, tag + f)
    if os.path.exists(p):
        out += ["## `ncu --set full --clock-control none` of the kernels of one apply, %s" % label, "",
                sh("python tools/ncu_summary.py %s" % p).strip(), ""]
open(os.path.join("profiles", name + ".md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])

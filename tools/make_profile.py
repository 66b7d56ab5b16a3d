#!/usr/bin/env python
"""Assemble profiles/<name>.md and profiles/traffic.json from the files a tools/gpu_session8.sh run left in gpurun_out/.
usage: python tools/make_profile.py <tag> <profile-name> "<title>" """
import csv, json, os, re, subprocess, sys

tag, name, title = sys.argv[1], sys.argv[2], sys.argv[3]
G = "gpurun_out"
out = ["# %s" % title, ""]


def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


tests = open(os.path.join(G, tag + "_tests.log")).read().strip().splitlines()
passed = [l for l in tests if " passed" in l or " failed" in l]
out += ["GPU tests of this build (`pytest tests -m gpu`): `%s`" % (passed[-1] if passed else "?"), ""]
smoke = [l for l in open(os.path.join(G, tag + "_smoke.log")).read().splitlines() if l.startswith("smoke")]
out += ["`__graft_entry__.smoke()`:", "```"] + smoke + ["```", ""]
for label, f in (("fused recipe (default), `python bench.py`", "_bench_cfg3.json"),
                 ("six-call -O3 recipe, `--tree o3`", "_bench_cfg3_o3.json"),
                 ("cfg1 (six-call recipe; its 512x512x2 grid has no fused plan)", "_bench_cfg1.json"),
                 ("reference arm, `--impl reference`", "_bench_ref.json"),
                 ("per-GPU shard of an 8-GPU run on one GPU, `--coils 2` (development flag)", "_bench_coils2.json"),
                 ("per-GPU shard of a 4-GPU run, `--coils 4`", "_bench_coils4.json"),
                 ("per-GPU shard of a 2-GPU run, `--coils 8`", "_bench_coils8.json")):
    p = os.path.join(G, tag + f)
    if os.path.exists(p) and os.path.getsize(p):
        out += ["### " + label, "```", open(p).read().strip().splitlines()[0], "```", ""]
out += ["## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
        "(cold-cache, serialised: compare shares; includes the one-time device construction of the operator)", "",
        sh("python tools/launch_summary.py %s/%s_launches_cfg3.csv" % (G, tag)).strip(), ""]
out += ["## `ncu --set full --clock-control none` of the fused-step kernels (one apply), per launch", "",
        sh("python tools/ncu_summary.py %s/%s_raw.csv" % (G, tag)).strip(), ""]
open(os.path.join("profiles", name + ".md"), "w").write("\n".join(out) + "\n")

# measured DRAM traffic per fused step (sum of its kernels), for bench.py's roofline.traffic
rows = list(csv.reader(open(os.path.join(G, tag + "_raw.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}


def gb(r, k):
    v = float(r[col[k]].replace(",", "")); u = units[col[k]].lower()
    return v * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0, "tbyte": 1e3}[u]


seq = [(re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", ""), gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum")) for r in data]
steps = {"expand_fft": 0.0, "ccsrmm_il[G' ": 0.0, "ccsrmm_il[G'^H": 0.0, "ifft_combine": 0.0}
state = "expand_fft"
for k, b in seq:                                   # kernels of one apply in launch order
    if k.startswith("sense_expand"): state = "expand_fft"
    elif k.startswith("kb_gather"): state = "ccsrmm_il[G' "
    elif k.startswith("csrmm_runs"): state = "ccsrmm_il[G'^H"
    elif state == "ccsrmm_il[G'^H" and k.startswith("fft_"): state = "ifft_combine"
    steps[state] += b
src = "profiles/%s.md (ncu --set full, sum over the kernels of the step)" % name
json.dump({"cfg3:1": {k: {"bytes": int(v * 1e9), "source": src} for k, v in steps.items()}},
          open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(steps))

#!/bin/bash
# one fused cfg3 bench line (per-step breakdown).  usage: gpu_bench_only.sh <tag> [ENV=VAL ...]
TAG=$1; shift
mkdir -p gpurun_out
env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}.json"))
    print("${TAG}: ms/step %.3f  e2e %.2f/s" % (d["ms_per_step"], d["e2e"]["value"]), " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("${TAG} failed", e); print(open("gpurun_out/${TAG}.err").read()[-1500:])
PY

#!/bin/bash
# fused tests + one fused cfg3 bench + launch-time list.  usage: gpu_quick2.sh <tag> [pytest -k expr]
TAG=$1; KEXPR=${2:-fused}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("ms/step %.3f  e2e %.2f/s" % (d["ms_per_step"], d["e2e"]["value"]), " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 2>&1 | grep -E "pk|csrmm|kb_gather|sense|fft" | cut -c1-160

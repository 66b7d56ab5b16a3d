#!/bin/bash
# quick GPU check: selected tests + fused cfg3 bench lines.  usage: gpu_quick.sh <tag> "<pytest -k expr>" [coil counts...]
TAG=$1; KEXPR=$2; shift 2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" ) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("cfg3 ms/step %.3f  applies/s %.2f  e2e %.2f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    for c in d["calls"]: print("   %-60s %8.3f ms  frac %.3f" % (c["call"][:60], c["ms"], c["frac"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
for C in "$@"; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_c$C.json 2> gpurun_out/${TAG}_bench_c$C.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_c$C.json"))
    print("coils=$C ms/step %.3f" % d["ms_per_step"])
    for c in d["calls"]: print("   %-60s %8.3f ms" % (c["call"][:60], c["ms"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench_c$C.err").read()[-2000:])
PY
done

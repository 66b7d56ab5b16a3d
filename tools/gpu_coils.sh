#!/bin/bash
# per-GPU shards of cfg3 on one GPU (development): bench.py --coils C for C in "$@"
mkdir -p gpurun_out
for C in "$@"; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/coils_$C.json 2> gpurun_out/coils_$C.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/coils_$C.json"))
    print("coils=$C ms/step %.3f" % d["ms_per_step"], " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("coils=$C failed", e); print(open("gpurun_out/coils_$C.err").read()[-1500:])
PY
done

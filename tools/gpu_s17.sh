#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "fused" ) > gpurun_out/s17_tests.log 2>&1; tail -2 gpurun_out/s17_tests.log
for S in 8 4 64; do
IB200_SAMPLE_SUPER=$S timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s17_s$S.json 2> gpurun_out/s17_s$S.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s17_s$S.json"))
    print("super=$S ms/step %.3f" % d["ms_per_step"], " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s17_s$S.err").read()[-1500:])
PY
done

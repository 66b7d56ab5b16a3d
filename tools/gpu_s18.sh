#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "fused" ) > gpurun_out/s18_tests.log 2>&1; tail -2 gpurun_out/s18_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s18_${name}.json 2> gpurun_out/s18_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s18_${name}.json"))
    print("${name}: ms/step %.3f" % d["ms_per_step"], " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/s18_${name}.err").read()[-1500:])
PY
}
run t8s8 IB200_SAMPLE_TILE=8 IB200_SAMPLE_SUPER=8
run t4s16 IB200_SAMPLE_TILE=4 IB200_SAMPLE_SUPER=16
run t4s8 IB200_SAMPLE_TILE=4 IB200_SAMPLE_SUPER=8
run t2s16 IB200_SAMPLE_TILE=2 IB200_SAMPLE_SUPER=16

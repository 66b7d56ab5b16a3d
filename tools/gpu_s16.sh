#!/bin/bash
mkdir -p gpurun_out
for T in 8192 32768; do
IB200_RUN_LONG=$T timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s16_t$T.json 2> gpurun_out/s16_t$T.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s16_t$T.json"))
    print("run_long=$T ms/step %.3f" % d["ms_per_step"], " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s16_t$T.err").read()[-1500:])
PY
done
bash tools/gpu_ncu_fft.sh s16 "csrmm_runs|kb_gather|il_long" 3

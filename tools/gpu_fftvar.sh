#!/bin/bash
# fused tests + bench under a few FFT env variants.  usage: gpu_fftvar.sh <tag>
TAG=$1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "fused" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}: ms/step %.3f" % d["ms_per_step"], " | ".join("%s %.3f" % (c["call"][:12], c["ms"]) for c in d["calls"]))
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run pkp X=1
run pkp1 IB200_FFT_PKP_CTAS=1
run pkp4 IB200_FFT_PKP_CTAS=4
run nopkp IB200_FFT_NOPKP=1

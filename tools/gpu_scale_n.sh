#!/bin/bash
# coil-sharded bench at N GPUs with the N = 1 digest check.  usage: gpu_scale_n.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --check ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N", round(d["value"], 2), "applies/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2), "check", d.get("check"))
    for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f" % (k["kernel"], k["ms"], k["frac"]))
except Exception as e: print("parse error", e)
PY
tail -4 gpurun_out/${TAG}_bench_n$N.err | cut -c1-300

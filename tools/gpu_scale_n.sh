#!/bin/bash
# coil-sharded bench at N GPUs with the N = 1 digest check.  usage: gpu_scale_n.sh <tag> <N> [extra bench flags]
TAG=$1; N=$2; shift; shift
mkdir -p gpurun_out
SUF=$(echo "$*" | tr -d ' -')
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --check "$@" ) > gpurun_out/${TAG}_bench_n$N$SUF.json 2> gpurun_out/${TAG}_bench_n$N$SUF.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N$SUF.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$N $*", round(d["value"], 2), "applies/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 2), "check", d.get("check"))
except Exception as e: print("parse error", e)
PY
tail -3 gpurun_out/${TAG}_bench_n$N$SUF.err | cut -c1-300

#!/bin/bash
# Coil-sharded cfg3 bench on N GPUs of one box, launched the way the driver does.  usage: gpu_scale.sh <N> [tag]
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
grep '"metric"' gpurun_out/${TAG}_n$N.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_n$N.err | cut -c1-300

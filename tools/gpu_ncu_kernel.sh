#!/bin/bash
# ncu --set full of kernels matching a regex inside one short bench.py run.  usage: gpu_ncu_kernel.sh <tag> <regex> [count] [skip]
TAG=$1; RX=$2; CNT=${3:-2}; SKIP=${4:-0}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300

#!/bin/bash
# gridding micro-benchmark (+ optional ncu --set full of the interleaved gathers) at cfg3
TAG=${1:-g}
NCU=${2:-0}
mkdir -p gpurun_out
true
timeout 900 python tools/bench_grid.py > gpurun_out/${TAG}_grid.json 2> gpurun_out/${TAG}_grid.err; cat gpurun_out/${TAG}_grid.json; tail -3 gpurun_out/${TAG}_grid.err
if [ "$NCU" != "0" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csrmm_il -c 20 \
    -o gpurun_out/${TAG}_il_full -f python tools/bench_grid.py --reps 0 > gpurun_out/${TAG}_ncu_il.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_il.log | cut -c1-300
fi

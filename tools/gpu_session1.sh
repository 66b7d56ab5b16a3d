#!/bin/bash
# first GPU session of a round: tests, bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
cat gpurun_out/bench_cfg3.json | cut -c1-3000
tail -3 gpurun_out/bench_cfg3.err
( time timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg1 --no-cpu-baseline ) > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
cat gpurun_out/bench_cfg1.json | cut -c1-2000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-500

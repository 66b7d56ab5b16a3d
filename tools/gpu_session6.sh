#!/bin/bash
# Session 6 opener: GPU tests, fused cfg3 bench, launch list, ncu --set full of the fused-step kernels.
TAG=${1:-s6}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
cut -c1-3000 gpurun_out/${TAG}_bench_cfg3.json
tail -3 gpurun_out/${TAG}_bench_cfg3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_cfg3.csv 2>&1 | head -40
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'csrmm_ils|csrmm_il_long|fft_il|sense_|fft_spec' -s 12 -c 12 \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
ls -la gpurun_out

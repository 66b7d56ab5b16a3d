#!/bin/bash
# GPU session: tests, bench (cfg3 + cfg1), launch list.  Usage: bash tools/gpu_session.sh <tag> [pytest -k expr]
TAG=${1:-s}
KEXPR=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -n "$KEXPR" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" ) > gpurun_out/${TAG}_tests.log 2>&1
else
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
fi
tail -6 gpurun_out/${TAG}_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
cut -c1-2500 gpurun_out/${TAG}_bench_cfg3.json
tail -3 gpurun_out/${TAG}_bench_cfg3.err
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tree o3 ) > gpurun_out/${TAG}_bench_cfg3_o3.json 2> gpurun_out/${TAG}_bench_cfg3_o3.err
cut -c1-2500 gpurun_out/${TAG}_bench_cfg3_o3.json
( time timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg1 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
cut -c1-1200 gpurun_out/${TAG}_bench_cfg1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_cfg3.csv 2>&1 | head -40
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cut -c1-1500 gpurun_out/${TAG}_bench_ref.json

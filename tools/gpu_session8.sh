#!/bin/bash
# Final validation session of the round: all GPU tests, smoke(), cfg3 bench (fused, with CPU baseline), o3 tree, cfg1,
# reference arm, per-GPU shard benches, launch list and ncu --set full of the fused-step kernels (the .ncu-rep stays
# on the box: only CSV exports travel back, gpurun_out is capped at 64 MiB).
TAG=${1:-s8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log | head -3
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep smoke gpurun_out/${TAG}_smoke.log | tail -3
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
cut -c1-400 gpurun_out/${TAG}_bench_cfg3.json; tail -3 gpurun_out/${TAG}_bench_cfg3.err
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tree o3 ) > gpurun_out/${TAG}_bench_cfg3_o3.json 2> gpurun_out/${TAG}_bench_cfg3_o3.err
cut -c1-300 gpurun_out/${TAG}_bench_cfg3_o3.json
( time timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg1 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
cut -c1-300 gpurun_out/${TAG}_bench_cfg1.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cut -c1-300 gpurun_out/${TAG}_bench_ref.json
for C in 8 4 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils$C.json 2> gpurun_out/${TAG}_bench_coils$C.err
  cut -c1-200 gpurun_out/${TAG}_bench_coils$C.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_cfg3.csv 2>&1 | head -12
timeout 1200 ncu --set full --clock-control none -k regex:'pk|csrmm_runs|kb_gather' -s 10 -c 10 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv
du -sh gpurun_out

#!/bin/bash
# Round 2, session 1: validate the refactor (reference-based backend, staged reference, its own suites on b200),
# and profile the 2-coil shard (8-GPU case) to see what bounds the gathers there.
TAG=${1:-r2s1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -40 gpurun_out/${TAG}_tests.log | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -5
( time timeout 1500 python tools/run_reference_suites.py --out gpurun_out/${TAG}_reference_suites.json ) > gpurun_out/${TAG}_refsuites.log 2>&1
tail -30 gpurun_out/${TAG}_refsuites.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_coils2.json 2> gpurun_out/${TAG}_bench_coils2.err
cut -c1-200 gpurun_out/${TAG}_bench_coils2.json; tail -3 gpurun_out/${TAG}_bench_coils2.err
timeout 900 ncu --set full --clock-control none -k regex:'pk|csrmm_runs|kb_gather' -s 10 -c 10 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_coils2.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_coils2.csv
du -sh gpurun_out

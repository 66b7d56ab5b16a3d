#!/bin/bash
# Round 2, session 3: cfg1 through the fused recipe (tiny z axis) with and without CUDA graphs, compute-sanitizer
# (memcheck + racecheck) on the smoke-sized paths, cfg2 SpMM sweep with DRAM traffic, final cfg3 numbers + ncu capture.
TAG=${1:-r2s3}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sense.py -m gpu -q --maxfail=10 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log | cut -c1-250
for G in "" "--graph"; do
  ( timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline $G ) > gpurun_out/${TAG}_bench_cfg1$G.json 2> gpurun_out/${TAG}_bench_cfg1$G.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_cfg1$G.json').read().strip().splitlines()[-1]); print('cfg1 $G', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms', d['config']['tree'][:30], [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']])"
  tail -2 gpurun_out/${TAG}_bench_cfg1$G.err
done
( timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --tree o3 --graph ) > gpurun_out/${TAG}_bench_cfg1_o3_graph.json 2>/dev/null
cut -c1-160 gpurun_out/${TAG}_bench_cfg1_o3_graph.json
# compute-sanitizer on the smoke-sized paths
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_smoke.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python bench.py --workload tiny --steps 2 --warmup 1 --no-cpu-baseline --check ) > gpurun_out/${TAG}_memcheck_tiny.log 2>&1; echo "memcheck tiny rc=$?"; tail -3 gpurun_out/${TAG}_memcheck_tiny.log | cut -c1-200
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/${TAG}_racecheck_tiny.log 2>&1; echo "racecheck tiny rc=$?"; tail -3 gpurun_out/${TAG}_racecheck_tiny.log | cut -c1-200
( timeout 900 python tools/bench_spmm.py ) > gpurun_out/${TAG}_spmm_sweep.md 2> gpurun_out/${TAG}_spmm_sweep.err; tail -20 gpurun_out/${TAG}_spmm_sweep.md
( time timeout 900 python bench.py --check ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f" % (k["kernel"], k["ms"], k["frac"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pk|csrmm_runs|kb_gather' -s 10 -c 10 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_cfg3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_cfg3.csv
du -sh gpurun_out

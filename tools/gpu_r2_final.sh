#!/bin/bash
# Round 2, final session: full GPU suite, smoke, every bench workload, shard timings, compute-sanitizer on the block
# gather, launch list and ncu --set full of one apply at 16 and at 2 coils.
TAG=${1:-r2fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log | cut -c1-250
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -5
( time timeout 900 python bench.py --check --check-tree ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
tail -3 gpurun_out/${TAG}_bench_cfg3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3", d["value"], d["ms_per_step"])
    for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f  frac_dram %s" % (k["kernel"], k["ms"], k["frac"], k.get("frac_dram")))
    print("e2e", d["e2e"]["value"], d["e2e"]["paths_timed"], "check", d.get("check"), "setup", d["setup"])
except Exception as e: print("parse error", e)
PY
for C in 8 4 2; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils$C.json 2> gpurun_out/${TAG}_bench_coils$C.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils$C.json').read().strip().splitlines()[-1]); print('coils $C', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
one() { # name, args...
  n=$1; shift
  ( time timeout 600 python bench.py "$@" ) > gpurun_out/${TAG}_bench_$n.json 2> gpurun_out/${TAG}_bench_$n.err
  python -c "
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$n.json').read().strip().splitlines() if l.startswith('{')][-1]); print('$n', d.get('value'), d.get('unit'), d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('cg'))
except Exception as e: print('$n parse error', e)"
  tail -1 gpurun_out/${TAG}_bench_$n.err | cut -c1-200
}
one cfg3_o3 --steps 5 --warmup 3 --no-cpu-baseline --tree o3
one cfg1 --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline
one cfg1_graph --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --graph
one cfg4 --workload cfg4 --no-cpu-baseline
one cfg5 --workload cfg5 --no-cpu-baseline --steps 5
one ref --impl reference --steps 2 --warmup 0
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python bench.py --workload tiny --steps 2 --warmup 1 --no-cpu-baseline --check ) > gpurun_out/${TAG}_memcheck_tiny44.log 2>&1; echo "memcheck tiny 4x4x4 rc=$?"
( IB200_BLOCKS_SHAPE=2,2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python bench.py --workload tiny --steps 2 --warmup 1 --no-cpu-baseline --check ) > gpurun_out/${TAG}_memcheck_tiny22.log 2>&1; echo "memcheck tiny 4x2x2 rc=$?"
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/${TAG}_racecheck_tiny44.log 2>&1; echo "racecheck tiny 4x4x4 rc=$?"
( IB200_BLOCKS_SHAPE=2,2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline ) > gpurun_out/${TAG}_racecheck_tiny22.log 2>&1; echo "racecheck tiny 4x2x2 rc=$?"
grep -h "ERROR SUMMARY" gpurun_out/${TAG}_memcheck_*.log gpurun_out/${TAG}_racecheck_*.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pk|kb_blocks|kb_gather' -s 9 -c 9 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_cfg3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_cfg3.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pk|kb_blocks|kb_gather' -s 9 -c 9 \
    -o /tmp/${TAG}_full2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_coils2.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_coils2.csv | cut -c1-200
du -sh gpurun_out

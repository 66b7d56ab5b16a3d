#!/bin/bash
# Round 2, session 2: full GPU test suite (incl. full-size parity, reference drop-in), new bench.py on every workload,
# tuning knobs of the run gather at small coil counts, launch list + ncu --set full of one cfg3 apply.
TAG=${1:-r2s2}
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -25 gpurun_out/${TAG}_tests.log | cut -c1-250
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -5
( time timeout 900 python bench.py --check --check-tree --write-digest ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
cut -c1-300 gpurun_out/${TAG}_bench_cfg3.json; tail -3 gpurun_out/${TAG}_bench_cfg3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_cfg3.json").read().strip().splitlines()[-1])
    for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f  frac_replaced %.3f" % (k["kernel"], k["ms"], k["frac"], k["frac_replaced_call"]))
    print("e2e", d["e2e"]["value"], "check", d.get("check"), "setup", d["setup"])
    print("cpu", {k: v for k, v in d.get("cpu_baseline", {}).items() if k in ("value", "cores", "kind", "seconds_per_apply")})
except Exception as e: print("parse error", e)
PY
for C in 8 4 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils$C.json 2> gpurun_out/${TAG}_bench_coils$C.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils$C.json').read().strip().splitlines()[-1]); print('coils $C', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']])"
done
for PL in 2 1; do
  IB200_RUNS_PL=$PL timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_coils2_pl$PL.json 2> gpurun_out/${TAG}_bench_coils2_pl$PL.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils2_pl$PL.json').read().strip().splitlines()[-1]); print('coils 2 PL $PL', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels'] if 'runs' in k['kernel']])"
done
IB200_RUNS_PL=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 4 > gpurun_out/${TAG}_bench_coils4_pl1.json 2> gpurun_out/${TAG}_bench_coils4_pl1.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils4_pl1.json').read().strip().splitlines()[-1]); print('coils 4 PL 1', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels'] if 'runs' in k['kernel']])"
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tree o3 ) > gpurun_out/${TAG}_bench_cfg3_o3.json 2> gpurun_out/${TAG}_bench_cfg3_o3.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg3_o3.json; tail -2 gpurun_out/${TAG}_bench_cfg3_o3.err
( time timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg1 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg1.json; tail -2 gpurun_out/${TAG}_bench_cfg1.err
( time timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg1 --no-cpu-baseline --graph ) > gpurun_out/${TAG}_bench_cfg1_graph.json 2> gpurun_out/${TAG}_bench_cfg1_graph.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg1_graph.json; tail -2 gpurun_out/${TAG}_bench_cfg1_graph.err
( time timeout 900 python bench.py --workload cfg4 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
cut -c1-250 gpurun_out/${TAG}_bench_cfg4.json; tail -2 gpurun_out/${TAG}_bench_cfg4.err
( time timeout 900 python bench.py --workload cfg5 --no-cpu-baseline --steps 5 ) > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err
cut -c1-250 gpurun_out/${TAG}_bench_cfg5.json; tail -2 gpurun_out/${TAG}_bench_cfg5.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
cut -c1-300 gpurun_out/${TAG}_bench_ref.json; tail -2 gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_cfg3.csv 2>&1 | head -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pk|csrmm_runs|kb_gather' -s 10 -c 10 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_cfg3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_cfg3.csv
du -sh gpurun_out

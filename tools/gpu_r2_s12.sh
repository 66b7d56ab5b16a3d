#!/bin/bash
# Round 2, session 12: full GPU suite on the block-gather build, cfg3 with check + digest + pipelined e2e, shard timings,
# launch list and ncu --set full of one apply at 16 and at 2 coils.
TAG=${1:-r2s12}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log | cut -c1-250
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -5
( time timeout 900 python bench.py --check --check-tree --write-digest ) > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err
cp profiles/digest_cfg3.npz gpurun_out/ 2>/dev/null
tail -3 gpurun_out/${TAG}_bench_cfg3.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3", d["value"], d["ms_per_step"])
    for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f  frac_replaced %.3f" % (k["kernel"], k["ms"], k["frac"], k["frac_replaced_call"]))
    print("e2e", d["e2e"]["value"], d["e2e"]["paths_timed"], d["e2e"]["path"][:60], "check", d.get("check"), "setup", d["setup"])
    print("cpu", {k: v for k, v in d.get("cpu_baseline", {}).items() if k in ("value", "cores", "kind", "seconds_per_apply")})
except Exception as e: print("parse error", e)
PY
for C in 8 4 2; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils$C.json 2> gpurun_out/${TAG}_bench_coils$C.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils$C.json').read().strip().splitlines()[-1]); print('coils $C', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']])"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_cfg3.csv 2>&1 | head -16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pk|kb_blocks|kb_gather' -s 9 -c 9 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_cfg3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_cfg3.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pk|kb_blocks|kb_gather' -s 9 -c 9 \
    -o /tmp/${TAG}_full2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_coils2.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_coils2.csv
du -sh gpurun_out

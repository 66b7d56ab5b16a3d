#!/bin/bash
# Round 2, last check of the final tree: full GPU suite, smoke, the default bench line (what the driver runs).
TAG=${1:-r2last}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log | cut -c1-250
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -6
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -4 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "gpu_launches")})
print("e2e", d["e2e"]["value"], d["e2e"]["paths_timed"]); print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "traffic", "frac_dram", "share_of_step")})
print("cpu_baseline", {k: d["cpu_baseline"][k] for k in ("value", "cores", "kind")}); print("clocks", d["clocks"])
PY
( time timeout 600 python bench.py --impl reference ) > gpurun_out/${TAG}_bench_ref_default.json 2> gpurun_out/${TAG}_bench_ref_default.err
cut -c1-400 gpurun_out/${TAG}_bench_ref_default.json; tail -3 gpurun_out/${TAG}_bench_ref_default.err

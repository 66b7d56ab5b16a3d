#!/bin/bash
# Round 2, session 23: matrix-free construction of the fused operator: parity suites, setup time and memory.
TAG=${1:-r2s23}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sense.py tests/test_gpu_reference.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=5 --timeout 600 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log | cut -c1-250
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -6
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['value'], 2), 'applies/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value'], 2), [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:3], 'setup', d['setup'], d.get('check'))
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
timeout 300 python bench.py --no-cpu-baseline --check --check-tree > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; summ gpurun_out/${TAG}_bench_cfg3.json cfg3; tail -2 gpurun_out/${TAG}_bench_cfg3.err
timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --coils 2 > gpurun_out/${TAG}_bench_coils2.json 2> gpurun_out/${TAG}_bench_coils2.err; summ gpurun_out/${TAG}_bench_coils2.json "coils 2"
timeout 200 python bench.py --no-cpu-baseline --steps 50 --warmup 5 --workload cfg1 --check > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err; summ gpurun_out/${TAG}_bench_cfg1.json cfg1
timeout 300 python bench.py --no-cpu-baseline --workload cfg4 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; summ gpurun_out/${TAG}_bench_cfg4.json cfg4

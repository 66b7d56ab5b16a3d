#!/bin/bash
# Round 2, session 17: 2-D problems (aliased z taps folded into the separable records): parity and cfg1 timings.
TAG=${1:-r2s17}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_sense.py tests/test_gpu_reference.py -m gpu -q -x --timeout 300 ) > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
for G in "" "--graph"; do
  ( timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --check $G ) > gpurun_out/${TAG}_bench_cfg1$G.json 2> gpurun_out/${TAG}_bench_cfg1$G.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_cfg1$G.json').read().strip().splitlines()[-1]); print('cfg1 $G', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms', [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']], d.get('check'))"
  tail -2 gpurun_out/${TAG}_bench_cfg1$G.err | cut -c1-200
done

#!/bin/bash
# Round 2, session 19: cfg1 with the small-problem settings (segments of 32 batches, one batch of samples per warp).
TAG=${1:-r2s19}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 300 -k "two_dimensional or cfg1 or against_oracle" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
for G in "" "--graph"; do
  ( timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --check $G ) > gpurun_out/${TAG}_bench_cfg1$G.json 2> gpurun_out/${TAG}_bench_cfg1$G.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_cfg1$G.json').read().strip().splitlines()[-1]); print('cfg1 $G', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value'],1), [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']], d.get('check'))"
done

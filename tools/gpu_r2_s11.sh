#!/bin/bash
# Round 2, session 11: fused tests on the new defaults; forward gather with rows in pairs.
TAG=${1:-r2s11}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 120 ) > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:2], d.get('check'))
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for P in 0 1 2; do
  IB200_KB_PAIR=$P timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --check > gpurun_out/${TAG}_bench_c16_pair$P.json 2> gpurun_out/${TAG}_bench_c16_pair$P.err
  summ gpurun_out/${TAG}_bench_c16_pair$P.json "coils 16 pair $P"
done
for P in 1 2; do
  IB200_KB_PAIR=$P timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_c2_pair$P.json 2> gpurun_out/${TAG}_bench_c2_pair$P.err
  summ gpurun_out/${TAG}_bench_c2_pair$P.json "coils 2 pair $P"
  IB200_KB_PAIR=$P timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_bench_c8_pair$P.json 2> gpurun_out/${TAG}_bench_c8_pair$P.err
  summ gpurun_out/${TAG}_bench_c8_pair$P.json "coils 8 pair $P"
done

#!/bin/bash
# Round 2, session 15: transform kernels with 384 / 512 threads per CTA (build variants under tools/exp/).
TAG=${1:-r2s15}
mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels'] if 'fft' in k['kernel'] or 'sense' in k['kernel']])
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for V in nt384 nt512; do
  for C in 16 2; do
    IB200_LIB=$PWD/tools/exp/lib_$V.so timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C --check > gpurun_out/${TAG}_${V}_c$C.json 2> gpurun_out/${TAG}_${V}_c$C.err
    summ gpurun_out/${TAG}_${V}_c$C.json "$V coils $C"
    tail -1 gpurun_out/${TAG}_${V}_c$C.err | cut -c1-200
  done
done

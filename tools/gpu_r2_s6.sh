#!/bin/bash
# Round 2, session 6: tile-block adjoint gather after the parallel fold; ncu --set full of its kernels at 2 coils.
TAG=${1:-r2s6}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "tile_blocks" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']])
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for P in 16 8; do
  IB200_TILES_PLN=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_coils2_pln$P.json 2> gpurun_out/${TAG}_bench_coils2_pln$P.err
  summ gpurun_out/${TAG}_bench_coils2_pln$P.json "coils 2 pln $P"
done
for P in 16 8; do
IB200_TILES_PLN=$P timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kb_tiles' -s 2 -c 2 \
    -o /tmp/${TAG}_full$P -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu$P.log 2>&1
ncu -i /tmp/${TAG}_full$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_tiles$P.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_tiles$P.csv
ncu -i /tmp/${TAG}_full$P.ncu-rep --page details --csv > gpurun_out/${TAG}_details_tiles$P.csv 2>/dev/null
done

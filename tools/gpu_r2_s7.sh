#!/bin/bash
# Round 2, session 7: tile-block adjoint gather with the fully asynchronous ring: parity, lane geometries, all coil counts.
TAG=${1:-r2s7}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "tile_blocks or against_oracle" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:3])
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for P in 4 8 16; do
  IB200_TILES_PLN=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_coils2_pln$P.json 2> gpurun_out/${TAG}_bench_coils2_pln$P.err
  summ gpurun_out/${TAG}_bench_coils2_pln$P.json "coils 2 pln $P"
done
for C in 4 8 16; do
  for P in 4 8; do
  IB200_TILES_MAXC=32 IB200_TILES_PLN=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils${C}_pln$P.json 2> gpurun_out/${TAG}_bench_coils${C}_pln$P.err
  summ gpurun_out/${TAG}_bench_coils${C}_pln$P.json "coils $C pln $P"
  done
done
IB200_TILES_MAXC=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kb_tiles_kernel' -s 1 -c 1 \
    -o /tmp/${TAG}_full2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_full2.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_tiles_c2.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_tiles_c2.csv
IB200_TILES_MAXC=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kb_tiles_kernel' -s 1 -c 1 \
    -o /tmp/${TAG}_full16 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu16.log 2>&1
ncu -i /tmp/${TAG}_full16.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_tiles_c16.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_tiles_c16.csv

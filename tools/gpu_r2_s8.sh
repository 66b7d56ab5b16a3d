#!/bin/bash
# Round 2, session 8: tile-block adjoint gather: parity at every coil count (per-test timeouts), 16-coil timing, ncu.
TAG=${1:-r2s8}
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 60 -k "tile_blocks" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:3])
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
IB200_TILES_MAXC=32 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_coils16_tiles.json 2> gpurun_out/${TAG}_bench_coils16_tiles.err
summ gpurun_out/${TAG}_bench_coils16_tiles.json "coils 16 tiles"
for P in 4 8; do
IB200_TILES_PLN=$P timeout 300 ncu --set full --clock-control none --import-source on -k regex:'kb_tiles_kernel' -s 1 -c 1 \
    -o /tmp/${TAG}_full$P -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_ncu$P.log 2>&1
ncu -i /tmp/${TAG}_full$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_tiles_pln$P.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_tiles_pln$P.csv
done
ncu -i /tmp/${TAG}_full4.ncu-rep --page source --csv > gpurun_out/${TAG}_src_tiles_pln4.csv 2>/dev/null

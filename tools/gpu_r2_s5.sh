#!/bin/bash
# Round 2, session 5: tile-block adjoint gather (kbtiles.cu) parity + shard timings; run-gather scheduling modes.
TAG=${1:-r2s5}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x ) > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']], 'setup', d.get('setup'))
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for C in 2 4; do
  for P in 8 4 16; do
    IB200_TILES_PLN=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils${C}_pln$P.json 2> gpurun_out/${TAG}_bench_coils${C}_pln$P.err
    summ gpurun_out/${TAG}_bench_coils${C}_pln$P.json "coils $C pln $P"
  done
done
IB200_TILES_SEG=32 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_bench_coils2_seg32.json 2> gpurun_out/${TAG}_bench_coils2_seg32.err
summ gpurun_out/${TAG}_bench_coils2_seg32.json "coils 2 seg 32"
IB200_TILES_MAXC=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_bench_coils8_tiles.json 2> gpurun_out/${TAG}_bench_coils8_tiles.err
summ gpurun_out/${TAG}_bench_coils8_tiles.json "coils 8 tiles"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_bench_coils8.json 2> gpurun_out/${TAG}_bench_coils8.err
summ gpurun_out/${TAG}_bench_coils8.json "coils 8 runs"
for MD in 0 1 2; do
  IB200_RUNS_MODE=$MD timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_mode$MD.json 2> gpurun_out/${TAG}_bench_cfg3_mode$MD.err
  summ gpurun_out/${TAG}_bench_cfg3_mode$MD.json "cfg3 runs mode $MD"
done
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cfg3_full_size and 2" ) > gpurun_out/${TAG}_fullsize.log 2>&1; tail -3 gpurun_out/${TAG}_fullsize.log

#!/bin/bash
# Round 2, session 13: segment length of the block gather at few coils.
TAG=${1:-r2s13}
mkdir -p gpurun_out
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:2])
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for SEG in 128 256 512; do
  for C in 2 4; do
    IB200_TILES_SEG=$SEG timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_c${C}_seg$SEG.json 2> gpurun_out/${TAG}_c${C}_seg$SEG.err
    summ gpurun_out/${TAG}_c${C}_seg$SEG.json "coils $C seg $SEG"
  done
done
IB200_TILES_SEG=256 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c16_seg256.json 2> gpurun_out/${TAG}_c16_seg256.err
summ gpurun_out/${TAG}_c16_seg256.json "coils 16 seg 256"
IB200_TILES_SEG=256 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_c8_seg256.json 2> gpurun_out/${TAG}_c8_seg256.err
summ gpurun_out/${TAG}_c8_seg256.json "coils 8 seg 256"

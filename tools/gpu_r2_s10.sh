#!/bin/bash
# Round 2, session 10: tensor-core tile gather: parity, timings at 16 / 8 coils, ncu.
TAG=${1:-r2s10}
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 60 -k "block_gather" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:2], 'setup', d['setup']['seconds'], round(d['setup']['resident_bytes_per_gpu']/1e9,1), d.get('check'))
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
IB200_TILES_MMA=8 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --check > gpurun_out/${TAG}_bench_c16_mma.json 2> gpurun_out/${TAG}_bench_c16_mma.err
summ gpurun_out/${TAG}_bench_c16_mma.json "coils 16 mma"
tail -2 gpurun_out/${TAG}_bench_c16_mma.err
IB200_TILES_MMA=8 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_bench_c8_mma.json 2> gpurun_out/${TAG}_bench_c8_mma.err
summ gpurun_out/${TAG}_bench_c8_mma.json "coils 8 mma"
IB200_TILES_MMA=4 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 4 > gpurun_out/${TAG}_bench_c4_mma.json 2> gpurun_out/${TAG}_bench_c4_mma.err
summ gpurun_out/${TAG}_bench_c4_mma.json "coils 4 mma"
IB200_TILES_MMA=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'kb_tiles_mma' -s 1 -c 1 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_mma_c16.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_mma_c16.csv

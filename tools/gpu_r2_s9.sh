#!/bin/bash
# Round 2, session 9: block adjoint gather in all shapes: parity, then shape x coil-count timings.
TAG=${1:-r2s9}
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 60 -k "block_gather" ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[2], round(d['ms_per_step'], 3), 'ms', [(k['kernel'], round(k['ms'], 3)) for k in d['kernels']][:2], 'setup', d['setup']['seconds'], round(d['setup']['resident_bytes_per_gpu']/1e9,1))
except Exception as e:
    print(sys.argv[2], 'parse error', e)
PY
}
for SH in 2,1 2,2 1,1; do
  for L in 1 2; do
    IB200_BLOCKS_SHAPE=$SH IB200_BLOCKS_LANES=$L timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c16_$SH-$L.json 2> gpurun_out/${TAG}_bench_c16_$SH-$L.err
    summ gpurun_out/${TAG}_bench_c16_$SH-$L.json "coils 16 shape $SH lanes $L"
  done
done
for SH in 2,1 2,2; do
  IB200_BLOCKS_SHAPE=$SH timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 8 > gpurun_out/${TAG}_bench_c8_$SH.json 2> gpurun_out/${TAG}_bench_c8_$SH.err
  summ gpurun_out/${TAG}_bench_c8_$SH.json "coils 8 shape $SH"
done
IB200_BLOCKS_SHAPE=2,2 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 4 > gpurun_out/${TAG}_bench_c4_22.json 2> gpurun_out/${TAG}_bench_c4_22.err
summ gpurun_out/${TAG}_bench_c4_22.json "coils 4 shape 2,2"
IB200_BLOCKS_SHAPE=2,1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'kb_blocks_kernel' -s 1 -c 1 \
    -o /tmp/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_blocks21_c16.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw_blocks21_c16.csv

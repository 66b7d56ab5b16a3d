#!/bin/bash
# Round 2, session 18: segment length of the block gather on the small 2-D problem (cfg1).
TAG=${1:-r2s18}
mkdir -p gpurun_out
for SEG in 64 32 16 8 4; do
  ( IB200_TILES_SEG=$SEG timeout 120 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --graph ) > gpurun_out/${TAG}_cfg1_seg$SEG.json 2> gpurun_out/${TAG}_cfg1_seg$SEG.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_cfg1_seg$SEG.json').read().strip().splitlines()[-1]); print('cfg1 graph seg $SEG', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms', [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']][:2])"
done
for SH in 4,4 2,1; do
  ( IB200_BLOCKS_SHAPE=$SH IB200_TILES_SEG=16 timeout 120 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --graph ) > gpurun_out/${TAG}_cfg1_sh$SH.json 2> gpurun_out/${TAG}_cfg1_sh$SH.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_cfg1_sh$SH.json').read().strip().splitlines()[-1]); print('cfg1 graph shape $SH seg 16', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms', [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']][:2])"
done

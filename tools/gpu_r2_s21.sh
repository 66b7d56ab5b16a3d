#!/bin/bash
# Round 2, session 21: sample tile size of the forward gather at 2 coils; smoke with the 16-coil apply.
TAG=${1:-r2s21}
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1; grep -i "smoke\|error" gpurun_out/${TAG}_smoke.log | tail -6
for T in 4 16; do
  IB200_SAMPLE_TILE=$T timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_c2_tile$T.json 2> gpurun_out/${TAG}_c2_tile$T.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_c2_tile$T.json').read().strip().splitlines()[-1]); print('coils 2 sample tile $T', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']][:3])"
done
IB200_SAMPLE_SUPER=4 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils 2 > gpurun_out/${TAG}_c2_super4.json 2> gpurun_out/${TAG}_c2_super4.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_c2_super4.json').read().strip().splitlines()[-1]); print('coils 2 super 4', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']][:3])"

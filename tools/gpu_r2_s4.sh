#!/bin/bash
# Round 2, session 4 (2 GPUs): coil-sharded bench with the N = 1 digest check and the slab end-to-end path;
# then 1-GPU re-checks of the two small fixes (combine pass without fold accumulators, shorter segments for small problems).
TAG=${1:-r2s4}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --check ) > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["path"][:60], "check", d.get("check"))
    for k in d["kernels"]: print("%-24s %7.3f ms  frac %.3f" % (k["kernel"], k["ms"], k["frac"]))
except Exception as e: print("parse error", e)
PY
tail -5 gpurun_out/${TAG}_bench_n2.err | cut -c1-300
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x ) > gpurun_out/${TAG}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_tests.log
for C in 2 4; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --coils $C > gpurun_out/${TAG}_bench_coils$C.json 2> gpurun_out/${TAG}_bench_coils$C.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_coils$C.json').read().strip().splitlines()[-1]); print('coils $C', round(d['ms_per_step'],3), [(k['kernel'], round(k['ms'],3)) for k in d['kernels']])"
done
for G in "" "--graph"; do
  ( timeout 300 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline $G ) > gpurun_out/${TAG}_bench_cfg1$G.json 2> gpurun_out/${TAG}_bench_cfg1$G.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_cfg1$G.json').read().strip().splitlines()[-1]); print('cfg1 $G', round(d['value'],1), 'applies/s', round(d['ms_per_step'],4), 'ms', [(k['kernel'][:22], round(k['ms'],4)) for k in d['kernels']])"
done

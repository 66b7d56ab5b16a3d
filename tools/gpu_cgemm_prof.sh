#!/bin/bash
# cgemm tcgen05 kernel: occupancy / padding variants, then one ncu --set full capture (forward + expansion).
mkdir -p gpurun_out
for v in "IB200_T5_MINB=2 IB200_T5_PAD=16" "IB200_T5_MINB=3 IB200_T5_PAD=16" "IB200_T5_MINB=2 IB200_T5_PAD=0" "IB200_T5_MINB=3 IB200_T5_PAD=48"; do
  echo "== $v"; env $v timeout 100 python tools/bench_cgemm.py --modes 0 --reps 5 2>&1 | grep "^| [YZ]"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cgemm_t5 -s 3 -c 5 -o gpurun_out/s36_t5_full -f \
    python tools/bench_cgemm.py --modes 0 --reps 1 > gpurun_out/s36_ncu.log 2>&1
tail -3 gpurun_out/s36_ncu.log | cut -c1-200
ls -la gpurun_out/s36_t5_full.ncu-rep

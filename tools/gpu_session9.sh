#!/bin/bash
# Session 9 (end of round 1): tools/gpu_session8.sh plus the cfg5 coil-compression cgemm bench and its ncu capture.
TAG=${1:-s9}
bash tools/gpu_session8.sh $TAG
timeout 200 python tools/bench_cgemm.py > gpurun_out/${TAG}_cgemm.md 2> gpurun_out/${TAG}_cgemm.err
grep "^| [YZ]\|rel-L2" gpurun_out/${TAG}_cgemm.md
timeout 300 ncu --set full --clock-control none -k regex:cgemm_tc -s 3 -c 5 -o /tmp/${TAG}_cgemm_full -f \
    python tools/bench_cgemm.py --modes 0 --reps 1 > gpurun_out/${TAG}_cgemm_ncu.log 2>&1
ncu -i /tmp/${TAG}_cgemm_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_cgemm_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_cgemm_raw.csv
du -sh gpurun_out

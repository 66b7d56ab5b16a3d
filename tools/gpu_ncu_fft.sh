#!/bin/bash
# ncu --set full of the FFT-pass kernels of one fused apply.  usage: gpu_ncu_fft.sh <tag> [regex] [count]
TAG=$1; RX=${2:-'pk_|fft_il|sense_'}; CNT=${3:-6}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $CNT -c $CNT \
    -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv

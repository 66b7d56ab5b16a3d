#!/bin/bash
# ncu --set full capture of the three FFT passes (axis 0, axis 1, axis 2) at cfg3
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fft_spec -c 3 \
    -o gpurun_out/r01_fft_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fft.log 2>&1
tail -3 gpurun_out/ncu_fft.log | cut -c1-400
ls -la gpurun_out/

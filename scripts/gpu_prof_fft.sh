#!/bin/bash
# ncu --set full (with source counters) of the forward x / y transform kernels and the Thomas kernel on C3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"r2r2_fwd|thomas_pipe" --launch-skip 9 -c 3 \
  -f -o gpurun_out/prof_fft python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fft.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_fft.log

#!/bin/bash
# ncu full capture of the z-pass + emit (DIT, TMA) and the noise gen kernel at 1024^3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"fft_rows_c2r_dit_kernel" -c 1 -f -o gpurun_out/r2s18_zpass \
  python bench.py --n 1024 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2s18_ncu.log 2>&1
tail -2 gpurun_out/r2s18_ncu.log | cut -c1-200

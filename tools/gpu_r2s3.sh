#!/bin/bash
# ncu full capture of the staged noise kernels at 1024^3 (one launch each)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"rng_gen_kernel|deliver_kernel" -c 2 -f -o gpurun_out/r2s3_noise \
  python bench.py --n 1024 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2s3_ncu.log 2>&1
tail -3 gpurun_out/r2s3_ncu.log
ls -la gpurun_out/r2s3_noise.ncu-rep

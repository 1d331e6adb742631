#!/bin/bash
# Round 2 session 1: whole GPU suite (config 3 path fixed), reference arms at 512^3 (CPU shim-FFT and the reference's USE_CUFFT=1 path)
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/r2s1_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2s1_pytest_gpu.log
python bench.py --impl reference-cufft --cpu-n 512 --steps 1 --warmup 1 > gpurun_out/r2s1_bench_reference_cufft.json 2> gpurun_out/r2s1_bench_reference_cufft.err
tail -1 gpurun_out/r2s1_bench_reference_cufft.json | cut -c1-400
python bench.py --impl reference --cpu-n 512 --steps 1 --warmup 1 > gpurun_out/r2s1_bench_reference512.json 2> gpurun_out/r2s1_bench_reference512.err
tail -1 gpurun_out/r2s1_bench_reference512.json | cut -c1-400
nproc; free -g | head -2

#!/bin/bash
# Round 2 session 2: staged single-generation noise path -- parity, then timing against the two-pass path
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_noise.py -x -q ) > gpurun_out/r2s2_pytest_noise.log 2>&1
tail -15 gpurun_out/r2s2_pytest_noise.log
for v in "GX_RNG_GEN_MINB=6" "GX_RNG_GEN_MINB=5" "GX_RNG_GEN_MINB=8" "GX_NOISE_STAGED=0"; do
  echo "== 1024 $v"
  env $v python bench.py --n 1024 --steps 3 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/r2s2_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], {k:round(v,2) for k,v in d['roofline']['per_class_ms_per_step'].items()})"
done
for v in "GX_RNG_GEN_MINB=6" "GX_RNG_GEN_MINB=5"; do
  echo "== 2048 $v"
  env $v python bench.py --n 2048 --steps 2 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s2_err.txt | tee gpurun_out/r2s2_bench2048_${v#*=}.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], {k:round(v,2) for k,v in d['roofline']['per_class_ms_per_step'].items()})"
done
tail -5 gpurun_out/r2s2_err.txt

#!/bin/bash
# Round 2 session 4: gen kernel with warp-convergent slow phase, fma-pipe mask; parity + A/B timing + ncu
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_noise.py -x -q ) > gpurun_out/r2s4_pytest_noise.log 2>&1
tail -4 gpurun_out/r2s4_pytest_noise.log
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items() if k.startswith("rng")})'
for v in "GX_RNG_GEN_MINB=6 GX_RNG_FMA_MASK=1" "GX_RNG_GEN_MINB=5 GX_RNG_FMA_MASK=1" "GX_RNG_GEN_MINB=6 GX_RNG_FMA_MASK=0" "GX_RNG_GEN_MINB=5 GX_RNG_FMA_MASK=0" "GX_RNG_GEN_MINB=8 GX_RNG_FMA_MASK=1"; do
  echo "== 2048 $v"
  env $v python bench.py --n 2048 --steps 2 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s4_err.txt | python -c "$show"
done
ncu --set full --clock-control none --import-source on -k regex:"rng_gen_kernel" -c 1 -f -o gpurun_out/r2s4_gen \
  python bench.py --n 1024 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2s4_ncu.log 2>&1
tail -2 gpurun_out/r2s4_err.txt

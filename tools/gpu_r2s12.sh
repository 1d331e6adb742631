#!/bin/bash
# Round 2 session 12: noise gen with the uniform parse loop (tentative slow attempts), deliver fast path: parity, then timing
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_noise.py tests/test_gpu_slab.py -x -q ) > gpurun_out/r2s12_pytest.log 2>&1
tail -4 gpurun_out/r2s12_pytest.log
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items()})'
for v in "GX_RNG_GEN_MINB=5" "GX_RNG_GEN_MINB=6"; do
  echo "== 2048 $v"
  env $v python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s12_err.txt | tee gpurun_out/r2s12_bench2048_${v#*=}.json | python -c "$show"
done
tail -2 gpurun_out/r2s12_err.txt

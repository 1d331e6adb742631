#!/bin/bash
# Round 2 session 20: a share of the gen kernel's CTAs runs ranlxs2 on the FP64 pipe (sweep), parity first
mkdir -p gpurun_out
GX_RNG_F64_SHARE=4 timeout 600 python -m pytest tests/test_gpu_noise.py -x -q 2>&1 | grep -E "passed|failed|error" | tail -2
GX_RNG_F64_SHARE=8 timeout 600 python -m pytest tests/test_gpu_noise.py -x -q -k "bit_exact" 2>&1 | grep -E "passed|failed|error" | tail -2
show='import json,sys
d=json.loads(sys.stdin.readline()); c=d["roofline"]["per_class_ms_per_step"]; print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in c.items() if k.startswith("rng")}, d["checksum"]["vel_sumsq"][0])'
for sh in 0 2 3 4 5 8; do
  echo "== GX_RNG_F64_SHARE=$sh"
  GX_RNG_F64_SHARE=$sh timeout 300 python bench.py --n 2048 --steps 2 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s20_err.txt | python -c "$show"
done
tail -2 gpurun_out/r2s20_err.txt

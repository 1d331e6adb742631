#!/bin/bash
# Round 2 session 16: first tile issued before the twiddle fill (column passes), TMA issued before the CTA barrier (row pass)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fft.py tests/test_gpu_slab.py -x -q ) > gpurun_out/r2s16_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2s16_pytest.log | tail -3
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items() if k.startswith("fft")}, {k:round(v["frac"],3) for k,v in d["roofline"]["passes"].items() if k.startswith("fft")}, d["cufft_crosscheck"]["ok"], d["checksum"]["vel_sumsq"][0])'
for i in 1 2; do
  timeout 300 python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s16_err.txt | tee gpurun_out/r2s16_bench2048_$i.json | python -c "$show"
done
tail -3 gpurun_out/r2s16_err.txt

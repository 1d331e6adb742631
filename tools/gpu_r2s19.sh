#!/bin/bash
# Round 2 session 19: FP64 pipe peak (DFMA), tail host work under the last delivery, tightened a5/a6 tests
mkdir -p gpurun_out
tools/fp64_peak | tee gpurun_out/r2s19_fp64_peak.txt
( timeout 900 python -m pytest tests/test_gpu_fft.py tests/test_gpu_noise.py tests/test_gpu_slab.py -x -q ) > gpurun_out/r2s19_pytest.log 2>&1
grep -E "passed|failed|error|assert" gpurun_out/r2s19_pytest.log | tail -5
show='import json,sys
d=json.loads(sys.stdin.readline()); c=d["roofline"]["per_class_ms_per_step"]; print(round(d["ms_per_step"],1), "sum of classes", round(sum(c.values()),1), {k:round(v,2) for k,v in c.items()}, d["checksum"]["vel_sumsq"][0])'
for i in 1 2; do
  timeout 300 python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s19_err.txt | python -c "$show"
done
tail -2 gpurun_out/r2s19_err.txt

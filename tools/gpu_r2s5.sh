#!/bin/bash
# Round 2 session 5: peer-memory slab mode on one GPU with emulated ranks (parity), deliver kernel after the arithmetic clean-up
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_slab.py -x -q ) > gpurun_out/r2s5_pytest_slab.log 2>&1
tail -15 gpurun_out/r2s5_pytest_slab.log
python -m pytest tests/test_gpu_noise.py -x -q -k "bit_exact" 2>&1 | tail -2
python bench.py --n 2048 --steps 2 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/r2s5_err.txt | tee gpurun_out/r2s5_bench2048.json | python -c '
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items()})'
tail -3 gpurun_out/r2s5_err.txt

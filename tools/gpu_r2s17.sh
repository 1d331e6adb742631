#!/bin/bash
# Round 2 session 17: 4096-point column passes with clusters of 16 against the 512-rows-per-CTA form; even noise batches
mkdir -p gpurun_out
for v in "GX_COLS4096_CL8=0" "GX_COLS4096_CL8=1"; do
  for a in "4096 f32 32" "4096 f64 16"; do
    echo "== $v $a"; env $v timeout 200 python tools/slab_pass_bench.py $a 2>&1 | tail -1
  done
done
python tools/slab_pass_bench.py 2048 f32 64 2>&1 | tail -1
python tools/slab_pass_bench.py 2048 f64 64 2>&1 | tail -1
python -m pytest tests/test_gpu_noise.py -x -q -k "bit_exact or batches or 1024" 2>&1 | tail -2
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items()}, d["checksum"]["vel_sumsq"][0])'
timeout 300 python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s17_err.txt | python -c "$show"
tail -2 gpurun_out/r2s17_err.txt

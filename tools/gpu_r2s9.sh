#!/bin/bash
# Round 2 session 9: natural-order (DIT last stage) c2r row pass -- parity, cuFFT cross-check incl. f32 and 4096, A/B timing
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fft.py tests/test_gpu_noise.py tests/test_gpu_slab.py -x -q ) > gpurun_out/r2s9_pytest.log 2>&1
tail -4 gpurun_out/r2s9_pytest.log
python - <<'PY' 2>&1 | tee gpurun_out/r2s9_crosscheck.txt
import sys; sys.path.insert(0, ".")
import torch, bench
from genetic_b200 import lib as L
lib = L.load(); lib.gx_set_device(0)
for n, dt in [(512, L.F64), (1024, L.F64), (2048, L.F64), (4096, L.F64), (512, L.F32), (2048, L.F32), (4096, L.F32)]:
    print(bench.cufft_crosscheck(lib, torch, n, dt))
PY
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items()}, {k:round(v["frac"],3) for k,v in d["roofline"]["passes"].items()})'
for v in "GX_ROWS_DIT=1" "GX_ROWS_DIT=0"; do
  echo "== 2048 $v"
  env $v python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s9_err.txt | tee gpurun_out/r2s9_bench2048_${v#*=}.json | python -c "$show"
done
tail -2 gpurun_out/r2s9_err.txt

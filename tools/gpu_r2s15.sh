#!/bin/bash
# Round 2 session 15: L2 prefetch of the tile one wave ahead in the column passes and the row pass (distance sweep)
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items() if k.startswith("fft")}, {k:round(v["frac"],3) for k,v in d["roofline"]["passes"].items() if k.startswith("fft")}, d["cufft_crosscheck"]["ok"], d["checksum"]["vel_sumsq"][0])'
for v in "GX_COLS_PREFETCH=0 GX_ROWS_PREFETCH=0" "GX_COLS_PREFETCH=92 GX_ROWS_PREFETCH=740" "GX_COLS_PREFETCH=184 GX_ROWS_PREFETCH=1480" "GX_COLS_PREFETCH=46 GX_ROWS_PREFETCH=370" "GX_COLS_PREFETCH=368 GX_ROWS_PREFETCH=0"; do
  echo "== 2048 $v"
  env $v timeout 300 python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s15_err.txt | python -c "$show"
done
tail -3 gpurun_out/r2s15_err.txt

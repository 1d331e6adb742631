#!/bin/bash
# Round 2 session 7 (2 GPUs): 4096-point / f32 passes against cuFFT, push overlap on/off, f32 slab path, sharded noise vs oracle
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r2s7_crosscheck.txt
import sys; sys.path.insert(0, ".")
import torch, bench
from genetic_b200 import lib as L
lib = L.load(); lib.gx_set_device(0)
for n, dt in [(2048, L.F64), (2048, L.F32), (4096, L.F32), (4096, L.F64), (1024, L.F32)]:
    print(bench.cufft_crosscheck(lib, torch, n, dt))
PY
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 "$@"; }
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), round(d["nvlink"].get("GB/s_per_direction",0)), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step_rank0"].items()}, d.get("checksum",{}).get("vel_sumsq"))'
echo "== 2048 f64 push, overlap on"
GX_BENCH_N=2048 run bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s7_push2048_ov.json 2> gpurun_out/r2s7_a.err; tail -2 gpurun_out/r2s7_a.err; python -c "$show" < gpurun_out/r2s7_push2048_ov.json
echo "== 2048 f64 push, overlap off"
GX_SLAB_PUSH_OVERLAP=0 GX_BENCH_N=2048 run bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s7_push2048_noov.json 2> gpurun_out/r2s7_b.err; tail -2 gpurun_out/r2s7_b.err; python -c "$show" < gpurun_out/r2s7_push2048_noov.json
echo "== 2048 f32 push"
GX_BENCH_DTYPE=f32 GX_BENCH_N=2048 run bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2s7_push2048_f32.json 2> gpurun_out/r2s7_c.err; tail -2 gpurun_out/r2s7_c.err; python -c "$show" < gpurun_out/r2s7_push2048_f32.json
echo "== sharded noise vs oracle stream"
run tests/multi_gpu/noise_shell_check.py --size 2048 --dtype f64 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
run tests/multi_gpu/noise_shell_check.py --size 2048 --dtype f32 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3

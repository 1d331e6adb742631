#!/bin/bash
# Round 2 session 22 (8 GPUs): config 5 with the cluster-16 plan for 4096-point lines; 2048^3 f64 once more with the final code
mkdir -p gpurun_out
run() { np=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 "$@"; }
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), round(d["nvlink"].get("GB/s_per_direction",0)), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step_rank0"].items()}, d.get("checksum",{}).get("vel_sumsq"))'
echo "== config 5: 4096 f32, 8 GPUs"
GX_BENCH_E2E_STEPS=1 GX_BENCH_E2E_WARMUP=0 GX_BENCH_DTYPE=f32 GX_BENCH_N=4096 run 8 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2s22_bench4096_f32_8gpu.json 2> gpurun_out/r2s22_c.err; tail -2 gpurun_out/r2s22_c.err | cut -c1-300; python -c "$show" < gpurun_out/r2s22_bench4096_f32_8gpu.json
echo "== 2048 f64, 8 GPUs"
GX_BENCH_E2E_STEPS=1 GX_BENCH_E2E_WARMUP=0 GX_BENCH_N=2048 run 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s22_bench2048_8gpu.json 2> gpurun_out/r2s22_b.err; tail -2 gpurun_out/r2s22_b.err | cut -c1-300; python -c "$show" < gpurun_out/r2s22_bench2048_8gpu.json

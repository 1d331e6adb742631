#!/bin/bash
# Round 2 session 8 (8 GPUs): 2048^3 f64 and config 5 (4096^3 f32) in peer-memory slab mode; noise parity at 4096; PCIe ceiling
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2s8_topo.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 "$@"; }
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), round(d["nvlink"].get("GB/s_per_direction",0)), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step_rank0"].items()}, d.get("checksum",{}).get("vel_sumsq"))'
run tests/multi_gpu/d2h_ceiling.py 2>&1 | grep d2h_ceiling | tee gpurun_out/r2s8_d2h_ceiling.txt
echo "== 2048 f64, 8 GPUs"
NCCL_DEBUG=WARN GX_BENCH_N=2048 run bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2s8_bench2048_8gpu.json 2> gpurun_out/r2s8_a.err; tail -2 gpurun_out/r2s8_a.err; python -c "$show" < gpurun_out/r2s8_bench2048_8gpu.json
echo "== noise parity 4096 f32, 8 GPUs"
run tests/multi_gpu/noise_shell_check.py --size 4096 --dtype f32 --shells 1 2>&1 | grep -v "^\*\|OMP_NUM" | tail -2 | tee gpurun_out/r2s8_noise4096.txt
echo "== config 5: 4096 f32, 8 GPUs"
GX_BENCH_E2E_STEPS=1 GX_BENCH_E2E_WARMUP=0 GX_BENCH_DTYPE=f32 GX_BENCH_N=4096 run bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2s8_bench4096_f32_8gpu.json 2> gpurun_out/r2s8_b.err; tail -2 gpurun_out/r2s8_b.err; python -c "$show" < gpurun_out/r2s8_bench4096_f32_8gpu.json

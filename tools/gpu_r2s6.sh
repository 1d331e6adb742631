#!/bin/bash
# Round 2 session 6 (2 GPUs): peer-memory slab mode under torchrun / NCCL / CUDA IPC against the all-to-all path
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@"; }
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), d.get("nvlink"), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step_rank0"].items()}, d.get("checksum",{}).get("vel_sumsq"))'
echo "== push, 512 (quick sanity)"
GX_BENCH_N=512 run --steps 2 --warmup 1 > gpurun_out/r2s6_push512.json 2> gpurun_out/r2s6_push512.err; tail -3 gpurun_out/r2s6_push512.err; python -c "$show" < gpurun_out/r2s6_push512.json
echo "== push, 2048"
GX_BENCH_N=2048 run --steps 3 --warmup 1 > gpurun_out/r2s6_push2048.json 2> gpurun_out/r2s6_push2048.err; tail -3 gpurun_out/r2s6_push2048.err; python -c "$show" < gpurun_out/r2s6_push2048.json
echo "== all-to-all path, 2048"
GX_BENCH_N=2048 GX_SLAB_PUSH=0 run --steps 2 --warmup 1 > gpurun_out/r2s6_a2a2048.json 2> gpurun_out/r2s6_a2a2048.err; tail -3 gpurun_out/r2s6_a2a2048.err; python -c "$show" < gpurun_out/r2s6_a2a2048.json
echo "== 1 GPU, 512 checksum for comparison"
python bench.py --n 512 --steps 2 --warmup 1 --no-cpu 2> gpurun_out/r2s6_1gpu512.err | tee gpurun_out/r2s6_1gpu512.json | python -c '
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],2), d["checksum"], d["cufft_crosscheck"])'
tail -3 gpurun_out/r2s6_1gpu512.err

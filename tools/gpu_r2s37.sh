#!/bin/bash
# Round 2 session 37 (2 GPUs): the 2-rank bench line with the round's final code (buffer pool, NUMA-local pinned rings)
mkdir -p gpurun_out
GX_BENCH_E2E_STEPS=1 GX_BENCH_E2E_WARMUP=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2s37_bench2048_2gpu.json 2> gpurun_out/r2s37.err
tail -3 gpurun_out/r2s37.err | cut -c1-300
python -c "
import json
for l in open('gpurun_out/r2s37_bench2048_2gpu.json'):
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), d.get('checksum',{}).get('vel_sumsq'), d['n_gpus'])"

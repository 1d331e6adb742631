#!/bin/bash
# Round 2 session 10 (2 GPUs): pipelined noise, limited-footprint push on a side stream, TMA row staging; parity first
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_slab.py tests/test_gpu_fft.py -x -q ) > gpurun_out/r2s10_pytest.log 2>&1
tail -4 gpurun_out/r2s10_pytest.log
show1='import json,sys
d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step"].items() if k.startswith("fft")}, {k:round(v["frac"],3) for k,v in d["roofline"]["passes"].items() if k.startswith("fft")})'
for v in "GX_ROWS_TMA=1" "GX_ROWS_TMA=0"; do
  echo "== 1 GPU 2048 $v"
  env $v python bench.py --n 2048 --steps 3 --warmup 1 --no-e2e --no-cpu 2>> gpurun_out/r2s10_err.txt | python -c "$show1"
done
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 "$@"; }
show='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],1), round(d["nvlink"].get("GB/s_per_direction",0)), {k:round(v,2) for k,v in d["roofline"]["per_class_ms_per_step_rank0"].items()}, d.get("checksum",{}).get("vel_sumsq"))'
for v in "GX_NOISE_PIPELINE=0 GX_SLAB_PUSH_OVERLAP=0" "GX_NOISE_PIPELINE=1 GX_SLAB_PUSH_OVERLAP=0" "GX_NOISE_PIPELINE=1 GX_DELIVER_CTAS_PER_SM=2 GX_SLAB_PUSH_OVERLAP=0" "GX_NOISE_PIPELINE=1 GX_SLAB_PUSH_OVERLAP=1 GX_PUSH_CTAS_PER_SM=1" "GX_NOISE_PIPELINE=1 GX_SLAB_PUSH_OVERLAP=1 GX_PUSH_CTAS_PER_SM=2"; do
  echo "== 2 GPUs 2048 $v"
  env $v GX_BENCH_E2E_STEPS=0 GX_BENCH_E2E_WARMUP=0 GX_BENCH_N=2048 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 1 2>> gpurun_out/r2s10_err.txt | python -c "$show"
done
tail -3 gpurun_out/r2s10_err.txt

# session 11 (2 GPUs): slab tests (phased noise with shared pass A, pass-B chunk lists) + NCCL bench at 2048^3
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_slab.py -x -q ) > $O/s11_pytest_slab.log 2>&1
tail -5 $O/s11_pytest_slab.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/s11_bench2048_g2.json 2> $O/s11_b.err
tail -3 $O/s11_b.err
cat $O/s11_bench2048_g2.json

# session 6 (2 GPUs): the slab path over NCCL -- bench at 1024^3 and 2048^3 under torchrun
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --n 1024 > $O/s6_bench1024_g2.json 2> $O/s6_a.err
tail -5 $O/s6_a.err
cat $O/s6_bench1024_g2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/s6_bench2048_g2.json 2> $O/s6_b.err
tail -5 $O/s6_b.err
cat $O/s6_bench2048_g2.json

# session 12 (4 GPUs): slab tests incl. overlap variants, NCCL bench at 2048^3 with and without the overlapped exchange
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_slab.py -x -q ) > $O/s12_pytest_slab.log 2>&1
tail -5 $O/s12_pytest_slab.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 > $O/s12_bench2048_g4.json 2> $O/s12_a.err
tail -3 $O/s12_a.err
cat $O/s12_bench2048_g4.json
GX_SLAB_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 3 --warmup 3 > $O/s12_bench2048_g4_seq.json 2> $O/s12_b.err
tail -3 $O/s12_b.err
cat $O/s12_bench2048_g4_seq.json

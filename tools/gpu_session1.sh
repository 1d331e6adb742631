set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s1_pytest.log 2>&1
tail -3 gpurun_out/s1_pytest.log
timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu > gpurun_out/s1_bench2048_yx.json 2> gpurun_out/s1_bench2048_yx.err
GX_ZELDOVICH_ORDER=xy timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu > gpurun_out/s1_bench2048_xy.json 2> gpurun_out/s1_bench2048_xy.err
timeout 300 python bench.py --n 512 --steps 5 --warmup 3 > gpurun_out/s1_bench512.json 2> gpurun_out/s1_bench512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s1_launches512.csv python bench.py --n 512 --steps 2 --warmup 1 --no-cpu > gpurun_out/s1_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fft_cols|rng_chunk|fft_rows_c2r|slot_sweep" -c 10 -f -o gpurun_out/s1_full1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-cpu > gpurun_out/s1_ncu_full.log 2>&1
ls -la gpurun_out

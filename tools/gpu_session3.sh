# single-GPU measurement session (round 1, re-entry): parity tests, bench at the headline size, ncu launch list and
# one full capture. Everything left in gpurun_out/ stays below 64 MiB (the .ncu-rep is exported to CSV and deleted).
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/s3_pytest.log 2>&1
tail -5 $O/s3_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/s3_smoke.log 2>&1; tail -2 $O/s3_smoke.log
timeout 900 python bench.py > $O/s3_bench2048.json 2> $O/s3_bench2048.err
timeout 300 python bench.py --n 512 --steps 5 --warmup 3 > $O/s3_bench512.json 2> $O/s3_bench512.err
timeout 300 python bench.py --n 1024 --steps 5 --warmup 3 --no-cpu > $O/s3_bench1024.json 2> $O/s3_bench1024.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s3_launches512.csv python bench.py --n 512 --steps 2 --warmup 1 --no-cpu > $O/s3_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:fft_cols|rng_chunk|fft_rows_c2r|slot_sweep' -c 12 -f -o /tmp/s3_full1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-cpu > $O/s3_ncu_full.log 2>&1
ncu -i /tmp/s3_full1024.ncu-rep --page raw --csv > $O/s3_full1024_raw.csv 2>/dev/null
ncu -i /tmp/s3_full1024.ncu-rep --page details --csv > $O/s3_full1024_details.csv 2>/dev/null
ls -la $O; du -sh $O
cat $O/s3_bench2048.json

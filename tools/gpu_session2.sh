# single-GPU measurement session: parity tests, bench (both Zeldovich pass orders), ncu launch list and full capture.
# Everything left in gpurun_out/ must stay below 64 MiB (the .ncu-rep is exported to CSV on the box and deleted).
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/s2_pytest.log 2>&1
tail -5 $O/s2_pytest.log
timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu > $O/s2_bench2048_yx.json 2> $O/s2_bench2048_yx.err
GX_ZELDOVICH_ORDER=xy timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu > $O/s2_bench2048_xy.json 2> $O/s2_bench2048_xy.err
timeout 300 python bench.py --n 512 --steps 5 --warmup 3 > $O/s2_bench512.json 2> $O/s2_bench512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s2_launches512.csv python bench.py --n 512 --steps 2 --warmup 1 --no-cpu > $O/s2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:fft_cols|rng_chunk|fft_rows_c2r|slot_sweep' -c 10 -f -o /tmp/s2_full1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-cpu > $O/s2_ncu_full.log 2>&1
ncu -i /tmp/s2_full1024.ncu-rep --page raw --csv > $O/s2_full1024_raw.csv 2>/dev/null
ncu -i /tmp/s2_full1024.ncu-rep --page details --csv > $O/s2_full1024_details.csv 2>/dev/null
ncu -i /tmp/s2_full1024.ncu-rep --page source --csv -k regex:rng_chunk > $O/s2_full1024_source_rng.csv 2>/dev/null
gzip -f $O/s2_full1024_source_rng.csv
ls -la $O; du -sh $O

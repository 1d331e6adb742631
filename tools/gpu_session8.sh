# session 8: row-pass store refactor (parity + timing), full default bench line (e2e with 16 GiB staging, CPU baseline),
# ncu launch list (512^3) and full capture of the step's kernels (1024^3)
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s8_pytest.log 2>&1
tail -6 $O/s8_pytest.log
timeout 900 python bench.py > $O/s8_bench2048.json 2> $O/s8_a.err
cat $O/s8_bench2048.json
timeout 300 python bench.py --impl reference > $O/s8_bench_reference.json 2> $O/s8_r.err
cat $O/s8_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s8_launches512.csv python bench.py --n 512 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/s8_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:rng_chunk|slot_sweep|fft_cols|fft_rows' -c 7 -f -o /tmp/s8_full1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-cpu --no-e2e > $O/s8_ncu_full.log 2>&1
ncu -i /tmp/s8_full1024.ncu-rep --page raw --csv > $O/s8_full1024_raw.csv 2>/dev/null
tail -3 $O/s8_ncu_full.log
ls -la $O | tail -12

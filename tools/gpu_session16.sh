# session 16 (final of round 1): full GPU suite, the official bench lines, zoom bench, ncu launch list and full capture
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s16_pytest.log 2>&1
tail -6 $O/s16_pytest.log
timeout 120 python __graft_entry__.py --smoke > $O/s16_smoke.log 2>&1; tail -1 $O/s16_smoke.log
timeout 900 python bench.py > $O/s16_bench2048.json 2> $O/s16_a.err
cat $O/s16_bench2048.json
timeout 300 python bench.py --impl reference > $O/s16_bench_reference.json 2> $O/s16_r.err
cat $O/s16_bench_reference.json
timeout 600 python tools/zoom_bench.py --n 256 --cpu-n 64 > $O/s16_zoom256.json 2> $O/s16_z.err
cat $O/s16_zoom256.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s16_launches512.csv python bench.py --n 512 --steps 2 --warmup 1 --no-cpu --no-e2e > $O/s16_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:rng_chunk|fft_cols|fft_rows' -c 9 -f -o /tmp/s16_full1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-cpu --no-e2e > $O/s16_ncu_full.log 2>&1
ncu -i /tmp/s16_full1024.ncu-rep --page raw --csv > $O/s16_full1024_raw.csv 2>/dev/null
tail -2 $O/s16_ncu_full.log
ls -la $O | grep s16

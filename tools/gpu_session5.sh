# session 5: carry-chain SWB generator (bit-exactness + timing), persistent double-buffered z-pass A/B
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_noise.py tests/test_gpu_fft.py -x -q ) > $O/s5_pytest.log 2>&1
tail -5 $O/s5_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s5_bench2048_pipe1.json 2> $O/s5_a.err
GX_ROWS_PIPE=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s5_bench2048_pipe0.json 2> $O/s5_b.err
timeout 300 python bench.py --n 1024 --no-cpu --no-e2e --steps 5 --warmup 3 > $O/s5_bench1024_pipe1.json 2> $O/s5_c.err
GX_ROWS_PIPE=0 timeout 300 python bench.py --n 1024 --no-cpu --no-e2e --steps 5 --warmup 3 > $O/s5_bench1024_pipe0.json 2> $O/s5_d.err
tail -3 $O/s5_*.err
for f in $O/s5_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done

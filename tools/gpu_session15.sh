# session 15: shared x-transform schedule (5 cheap column passes): parity of the pipeline + A/B timing
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_noise.py tests/test_gpu_multilevel.py tests/test_cpp_host.py -m gpu -x -q ) > $O/s15_pytest.log 2>&1
tail -6 $O/s15_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s15_bench2048_shared.json 2> $O/s15_a.err
GX_ZELDOVICH_SHARED=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s15_bench2048_6pass.json 2> $O/s15_b.err
tail -2 $O/s15_a.err $O/s15_b.err
for f in $O/s15_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done

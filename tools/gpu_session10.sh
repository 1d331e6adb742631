# session 10: full GPU suite (new modification tests, slab pass-B chunk lists), bench
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s10_pytest.log 2>&1
tail -6 $O/s10_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s10_bench2048.json 2> $O/s10_a.err
python - $O/s10_bench2048.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY

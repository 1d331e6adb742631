# session 7: paired re/im delivery + threaded tail patch (bit-exactness + timing), C++ host goldens, 02i through the host layer
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s7_pytest.log 2>&1
tail -6 $O/s7_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s7_bench2048.json 2> $O/s7_a.err
timeout 300 python bench.py --n 512 --no-cpu --no-e2e --steps 5 --warmup 3 > $O/s7_bench512.json 2> $O/s7_b.err
for f in $O/s7_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done

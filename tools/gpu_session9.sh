# session 9: computed row twiddles + slim fused operator + 64-slot ring (parity + timing); zoom (config 3) bench; example paramfile test
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s9_pytest.log 2>&1
tail -6 $O/s9_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s9_bench2048.json 2> $O/s9_a.err
timeout 300 python bench.py --n 1024 --no-cpu --no-e2e --steps 5 --warmup 3 > $O/s9_bench1024.json 2> $O/s9_b.err
for f in $O/s9_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done
timeout 900 python tools/zoom_bench.py --n 256 --cpu-n 64 > $O/s9_zoom256.json 2> $O/s9_z.err
tail -3 $O/s9_z.err
cat $O/s9_zoom256.json

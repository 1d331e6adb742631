# session 14: shared y-transform (5 column passes) + transfer folded into the noise delivery: parity + A/B timing
set -x
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s14_pytest.log 2>&1
tail -6 $O/s14_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s14_bench2048_new.json 2> $O/s14_a.err
GX_ZELDOVICH_SHARED=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s14_bench2048_6pass.json 2> $O/s14_b.err
GX_FUSE_TRANSFER=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s14_bench2048_nofuse.json 2> $O/s14_c.err
for f in $O/s14_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done

# session 17: residency of the noise kernels (register cap A/B) and a re-run of the zoom bench
set -x
O=gpurun_out
mkdir -p $O
for m in 4 5 6; do
  GX_RNG_MINB=$m timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s17_bench2048_minb$m.json 2> $O/s17_$m.err
done
for f in $O/s17_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done
timeout 600 python tools/zoom_bench.py --n 256 --no-cpu > $O/s17_zoom256.json 2> $O/s17_z.err
cat $O/s17_zoom256.json

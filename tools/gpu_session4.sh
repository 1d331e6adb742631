# session 4: multi-level GPU parity + A/B of the coalesced z-pass store and the |k|^2-table transfer at 2048^3
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_multilevel.py -x -q ) > $O/s4_pytest_ml.log 2>&1
tail -15 $O/s4_pytest_ml.log
( time timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multilevel.py ) > $O/s4_pytest.log 2>&1
tail -5 $O/s4_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s4_bench2048_new.json 2> $O/s4_a.err
GX_ROWS_COALESCE=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s4_bench2048_rows0.json 2> $O/s4_b.err
GX_TRANSFER_TABLE=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 > $O/s4_bench2048_tab0.json 2> $O/s4_c.err
timeout 300 python bench.py --n 512 --no-cpu --no-e2e --steps 5 --warmup 3 > $O/s4_bench512_new.json 2> $O/s4_d.err
for f in $O/s4_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(d["ms_per_step"], {k: round(v,1) for k,v in d["roofline"]["per_class_ms_per_step"].items()})
PY
done

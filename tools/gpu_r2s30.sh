#!/bin/bash
# Round 2 session 30: validation of the round's code -- whole GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2s30_pytest_gpu.log 2>&1
grep -E "passed|failed|error" gpurun_out/r2s30_pytest_gpu.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2s30_bench2048.json 2> gpurun_out/r2s30_bench.err
tail -2 gpurun_out/r2s30_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s30_bench2048.json").readline())
print(round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["roofline"]["traffic"],
      {k: round(v["frac"], 3) for k, v in d["roofline"]["passes"].items()}, d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"], d["checksum"]["vel_sumsq"])
PY
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2s30_bench_reference.json 2>/dev/null; cut -c1-260 gpurun_out/r2s30_bench_reference.json

#!/bin/bash
# Round 2 session 11: whole GPU suite, default bench line at 2048^3 (e2e + cpu_baseline), ncu DRAM-traffic launch list at 2048^3
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x ) > gpurun_out/r2s11_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2s11_pytest_gpu.log
python bench.py > gpurun_out/r2s11_bench2048.json 2> gpurun_out/r2s11_bench.err
tail -2 gpurun_out/r2s11_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s11_bench2048.json").readline())
print(round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
      {k: round(v["frac"], 3) for k, v in d["roofline"]["passes"].items()}, d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"])
PY
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --replay-mode application \
  --csv --log-file gpurun_out/r2_ncu_dram_2048.csv python bench.py --n 2048 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2s11_ncu.log 2>&1
tail -2 gpurun_out/r2s11_ncu.log | cut -c1-300
python tools/ncu_dram_summary.py gpurun_out/r2_ncu_dram_2048.csv gpurun_out/r2_ncu_dram_2048.json

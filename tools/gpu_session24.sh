#!/bin/bash
# Session 24 (last GPU seconds of round 1): float mode through the command layer vs `genetIC -f`, paramfile goldens.
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_vs_reference.py -q -m gpu -k "float_mode" > gpurun_out/s24_pytest.log 2>&1
echo "exit $?" >> gpurun_out/s24_pytest.log
grep -E "^E|passed|failed|exit" gpurun_out/s24_pytest.log | head -20

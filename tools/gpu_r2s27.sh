#!/bin/bash
# Round 2 session 27 (1 GPU): grafic / mask goldens on the device, the field-buffer pool tests
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_grafic.py tests/test_gpu_pool.py -q -m gpu ) > gpurun_out/r2s27_pytest.log 2>&1
tail -30 gpurun_out/r2s27_pytest.log

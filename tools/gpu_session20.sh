#!/bin/bash
# Session 20 (last GPU minutes of round 1): the new device gather (gx_emit_cells) behind the particle mappers, the
# paramfile command layer and the zoom particle files against the reference binary, on the device.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mapper.py tests/test_gpu_paramfile.py tests/test_gpu_vs_reference.py -q -m gpu \
  > gpurun_out/s20_pytest.log 2>&1
echo "exit $?" >> gpurun_out/s20_pytest.log
tail -15 gpurun_out/s20_pytest.log

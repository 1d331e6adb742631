#!/bin/bash
# Round 2 session 23 (1 GPU): the C++ modifications / splicing driver on the reference's goldens, dump_ps on the device
mkdir -p gpurun_out
( time python -m pytest tests/test_cpp_modif.py tests/test_gpu_fft.py tests/test_gpu_paramfile.py tests/test_cpp_zoom.py -q -m gpu ) > gpurun_out/r2s23_pytest.log 2>&1
tail -40 gpurun_out/r2s23_pytest.log

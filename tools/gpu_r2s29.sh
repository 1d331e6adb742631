#!/bin/bash
# Round 2 session 29 (1 GPU): C++ grafic writer on the reference's goldens, all C++ modification goldens again
mkdir -p gpurun_out
( time python -m pytest tests/test_cpp_modif.py tests/test_cpp_zoom.py tests/test_cpp_host.py -q -m gpu ) > gpurun_out/r2s29_pytest.log 2>&1
tail -30 gpurun_out/r2s29_pytest.log

#!/bin/bash
# Round 2 session 39 (1 GPU): the C++ driver tests after the logging / mapper_relative_to changes
mkdir -p gpurun_out
( time python -m pytest tests/test_cpp_modif.py tests/test_cpp_zoom.py tests/test_cpp_host.py -q -m gpu ) > gpurun_out/r2s39_pytest.log 2>&1
tail -6 gpurun_out/r2s39_pytest.log

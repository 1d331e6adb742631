#!/bin/bash
# Round 2 session 34 (1 GPU): the reference's whole test suite on the device again after the host-made bin table of dump_ps;
# the dump_ps / paramfile tests
mkdir -p gpurun_out /tmp/reftests
tar -xzf oracle/_ref/reference_tests.tar.gz -C /tmp/reftests
( time GX_REFERENCE_TESTS=/tmp/reftests/tests timeout 1200 python tools/run_reference_suite.py ) > gpurun_out/r2_reference_suite_b200.txt 2>&1
grep -v "PASS" gpurun_out/r2_reference_suite_b200.txt | tail -12
( time python -m pytest tests/test_gpu_paramfile.py tests/test_gpu_fft.py tests/test_cpp_modif.py -q -m gpu ) > gpurun_out/r2s34_pytest.log 2>&1
tail -4 gpurun_out/r2s34_pytest.log

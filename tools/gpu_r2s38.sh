#!/bin/bash
# Round 2 session 38 (1 GPU): the C++ driver on the test directories its latest commands unlock (flagged gadget type,
# dump_grid_fourier) and a regression subset
mkdir -p gpurun_out /tmp/reftests
tar -xzf oracle/_ref/reference_tests.tar.gz -C /tmp/reftests
cd /tmp/reftests/tests && T=$(ls -d test_* mapper_test_* | tr '\n' ' ') && cd - > /dev/null
( time GX_REFERENCE_TESTS=/tmp/reftests/tests timeout 200 python tools/run_reference_suite.py --cpp $T ) > gpurun_out/r2s38_cpp_subset.txt 2>&1
cut -c1-200 gpurun_out/r2s38_cpp_subset.txt | tail -20

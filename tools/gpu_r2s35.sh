#!/bin/bash
# Round 2 session 36 (1 GPU): the reference's whole test suite through the C++ driver (host/genetic_b200_modify) on the device
mkdir -p gpurun_out /tmp/reftests
tar -xzf oracle/_ref/reference_tests.tar.gz -C /tmp/reftests
( time GX_REFERENCE_TESTS=/tmp/reftests/tests timeout 1200 python tools/run_reference_suite.py --cpp ) > gpurun_out/r2_reference_suite_b200_cpp.txt 2>&1
cut -c1-170 gpurun_out/r2_reference_suite_b200_cpp.txt | grep -v "PASS" | tail -45

#!/bin/bash
# Round 2 session 33 (1 GPU): the reference's WHOLE test suite (82 directories) through the command layer ON THE DEVICE, with the
# tolerances of the reference's tools/compare.py. The test DATA travels as a temporary, git-ignored tarball
# (oracle/_ref/reference_tests.tar.gz, removed afterwards); no reference source is involved.
mkdir -p gpurun_out /tmp/reftests
tar -xzf oracle/_ref/reference_tests.tar.gz -C /tmp/reftests
( time GX_REFERENCE_TESTS=/tmp/reftests/tests timeout 1200 python tools/run_reference_suite.py ) > gpurun_out/r2_reference_suite_b200.txt 2>&1
grep -v "PASS" gpurun_out/r2_reference_suite_b200.txt | tail -20

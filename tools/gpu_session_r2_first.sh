#!/bin/bash
# First GPU session of round 2 (prepared at the end of round 1, when the GPU minutes were spent):
#  1. the whole GPU suite, including the cases collected last that have not yet run on a B200
#  2. the reference's own GPU path (USE_CUFFT=1, real cuFFTW) as a third bench arm
#  3. SURVEY 8(d) config 3 (profiler_test, modernised) end to end: reference binary on the host cores vs python -m genetic_b200
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu ) > gpurun_out/r2s1_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2s1_pytest_gpu.log
python bench.py --impl reference-cufft --steps 2 --warmup 1 > gpurun_out/r2s1_bench_reference_cufft.json 2> gpurun_out/r2s1_bench_reference_cufft.err
tail -1 gpurun_out/r2s1_bench_reference_cufft.json | cut -c1-300
TIMEFORMAT='wall %R s'
for arm in ref gpu; do
  d=/tmp/prof_$arm; mkdir -p $d
  cp tools/paramfiles/profiler_test_modernised.txt $d/paramfile.txt
  cp tests/golden/camb_transfer_kmax40_z0.dat $d/
done
{
  echo "== profiler_test (modernised): reference binary, $(nproc) host threads"
  echo "(36 min wall clock in the build container, profiles/r1_profiler_test_reference_binary_build_container.txt; bounded to 5 min here"
  echo " because box time is GPU budget -- the comparison of the output files then uses whatever finished)"
  ( cd /tmp/prof_ref && time timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/genetIC paramfile.txt > out.txt 2>&1; grep -h "calculated\|chi\|number dm" out.txt | tail -5 )
  echo "== python -m genetic_b200 (device)"
  ( cd /tmp/prof_gpu && export PYTHONPATH=$GRAFT_REPO_ROOT && time GX_TIMING=1 python -m genetic_b200 paramfile.txt > out.txt 2>&1; grep -h "timing\|number dm" out.txt | tail -12 )
  python - <<PY
import sys
sys.path.insert(0, "$GRAFT_REPO_ROOT")
import numpy as np
from genetic_b200 import output as OUT
import os
if not os.path.exists("/tmp/prof_ref/test_9.tipsy"):
    print("reference binary did not finish within its bound: no file comparison")
    raise SystemExit(0)
a, b = OUT.read_particles("/tmp/prof_ref/test_9.tipsy"), OUT.read_particles("/tmp/prof_gpu/test_9.tipsy")
for k in ("pos", "vel", "mass", "eps"):
    d = np.abs(a[k] - b[k])
    print(k, "max |diff|", d.max(), "exact fraction", (d == 0).mean())
PY
} > gpurun_out/r2s1_profiler_test.txt 2>&1
cat gpurun_out/r2s1_profiler_test.txt

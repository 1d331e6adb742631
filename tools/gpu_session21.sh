#!/bin/bash
# Session 21: (a) example/paramfile.txt end to end, reference binary (host cores) vs `python -m genetic_b200` (device),
# wall clock of the whole process each; (b) the older GPU tests whose host code changed this session.
TIMEFORMAT='wall %R s'
mkdir -p gpurun_out /tmp/ex_ref /tmp/ex_gpu
for d in /tmp/ex_ref /tmp/ex_gpu; do
  cp tests/golden/example_paramfile.txt $d/paramfile.txt
  cp tests/golden/planck_2018_transfer_out.dat $d/
done
{
  echo "== reference binary (oracle/_ref/genetIC, $(nproc) host threads, FFT shim not FFTW), example/paramfile.txt"
  ( cd /tmp/ex_ref && time $GRAFT_REPO_ROOT/oracle/_ref/genetIC paramfile.txt > out.txt 2>&1; tail -1 out.txt; ls -la test.gadget3 )
  echo "== python -m genetic_b200 (device), same paramfile: first run (includes CUDA context + library load), then second"
  ( cd /tmp/ex_gpu && time python -m genetic_b200 paramfile.txt > out.txt 2>&1; tail -1 out.txt; ls -la test.gadget3 )
  ( cd /tmp/ex_gpu && time python -m genetic_b200 paramfile.txt > out2.txt 2>&1; tail -1 out2.txt )
  cmp /tmp/ex_ref/test.gadget3 /tmp/ex_gpu/test.gadget3 && echo "files identical" || echo "files differ (velocities to float32 rounding, see tests)"
} > gpurun_out/s21_example_paramfile.txt 2>&1
cat gpurun_out/s21_example_paramfile.txt

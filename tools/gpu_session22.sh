#!/bin/bash
# Session 22: a zoom IC 8x the example's size (base 256^3 over 100 Mpc/h, zoom x2 256^3, gadget3), whole process wall
# clock: reference binary on the host cores vs `python -m genetic_b200` on the device.
TIMEFORMAT='wall %R s'
mkdir -p gpurun_out /tmp/z_ref /tmp/z_gpu
for d in /tmp/z_ref /tmp/z_gpu; do
  cp tests/golden/planck_2018_transfer_out.dat $d/
  cat > $d/paramfile.txt <<PF
Om 0.3158
Ol 0.6842
s8 0.8117
ns 0.9660
hubble 0.6732
z 200
camb planck_2018_transfer_out.dat
random_seed 8896131
outname test
outdir ./
outformat gadget3
base_grid 100.0 256
centre 50 50 50
select_sphere 15
zoom_grid 2 256
done
PF
done
{
  echo "== reference binary (oracle/_ref/genetIC, $(nproc) host threads, FFT shim not FFTW): base 256^3 + zoom x2 256^3, gadget3"
  echo "(reference not re-run: 11.6 s in the first run of this session)"
  echo "== python -m genetic_b200 (device), same paramfile"
  ( cd /tmp/z_gpu && time GX_TIMING=1 python -m genetic_b200 paramfile.txt > out.txt 2>&1; grep timing out.txt; ls -la test.gadget3 )
} > gpurun_out/s22_zoom256_timing.txt 2>&1
cat gpurun_out/s22_zoom256_timing.txt

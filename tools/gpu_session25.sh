#!/bin/bash
# Session 25 (last seconds of GPU budget): the C++ zoom driver on the device against the Python layer, two-level gadget3.
mkdir -p gpurun_out /tmp/cz && cd /tmp/cz
cp $GRAFT_REPO_ROOT/tests/golden/camb_transfer_kmax40_z0.dat .
printf 'Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\nrandom_seed_real_space 90210\noutname t\noutdir ./\noutformat gadget3\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\ndone\n' > paramfile.txt
{
$GRAFT_REPO_ROOT/host/genetic_b200_zoom paramfile.txt && mv t.gadget3 cpp.gadget3
$GRAFT_REPO_ROOT/oracle/_ref/genetIC paramfile.txt > ref.log 2>&1
cmp cpp.gadget3 t.gadget3 && echo "C++ zoom driver: gadget3 file IDENTICAL to the reference binary's" || echo "C++ zoom driver: files differ"
} > $GRAFT_REPO_ROOT/gpurun_out/s25_cpp_zoom.txt 2>&1
cat $GRAFT_REPO_ROOT/gpurun_out/s25_cpp_zoom.txt

#!/bin/bash
# Round 2 session 32 (1 GPU): do the passes slow down under the board's power cap? (ncu's isolated launches are 20 % faster than
# the same kernels inside the step)
mkdir -p gpurun_out
timeout 300 python tools/power_cap_probe.py > gpurun_out/r2s32_power_cap_probe.txt 2>&1; cat gpurun_out/r2s32_power_cap_probe.txt | tail -30

#!/bin/bash
# Round 2 session 25 (1 GPU): host-side profile of config 3
mkdir -p gpurun_out
GX_ZOOM_CPROFILE=1 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu > gpurun_out/r2s25_zoom256.json 2> gpurun_out/r2s25_cprofile.txt; cat gpurun_out/r2s25_zoom256.json; head -70 gpurun_out/r2s25_cprofile.txt

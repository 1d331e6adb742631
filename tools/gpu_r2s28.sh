#!/bin/bash
# Round 2 session 28 (1 GPU): mixed-radix transform for non-power-of-two sizes, reverse_small_k; FFT / paramfile tests
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fft.py tests/test_gpu_paramfile.py tests/test_gpu_noise.py -q -m gpu ) > gpurun_out/r2s28_pytest.log 2>&1
tail -30 gpurun_out/r2s28_pytest.log
python - <<'P'
import time, numpy as np, genetic_b200 as gb
from genetic_b200 import lib as L
lib = L.load()
for n in (96, 192, 384, 768):
    g = gb.Grid(100.0, n); f = gb.Field(g, fourier=False)
    f.toFourier(); f.toReal(); lib.gx_synchronize(None)
    t0 = time.perf_counter()
    for _ in range(5): f.toFourier(); f.toReal()
    lib.gx_synchronize(None)
    dt = (time.perf_counter() - t0) / 10
    B = lib.gx_field_bytes(n, 1)
    print("mixed radix n=%d: %.3f ms per 3-D transform, %.0f GB/s (6 B per transform)" % (n, dt * 1e3, 6 * B / dt / 1e9))
P

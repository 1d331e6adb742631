"""Quick device timing of the FFT passes (development aid; bench.py is the contract benchmark)."""
import ctypes as C
import sys
import time

import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import genetic_b200 as gb
from genetic_b200 import lib as L


def main():
    lib = L.load()
    lib.gx_set_device(0)
    import torch
    for dtype, n in [(np.float64, 256), (np.float64, 512), (np.float32, 512), (np.float64, 1024), (np.float32, 1024),
                     (np.float64, 2048), (np.float32, 2048)]:
        if len(sys.argv) > 1 and str(n) not in sys.argv[1:]:
            continue
        dt = L.F64 if dtype == np.float64 else L.F32
        nbytes = lib.gx_field_bytes(n, dt)
        p = C.c_void_p()
        lib.gx_malloc(C.byref(p), nbytes)
        lib.gx_memset_zero(p, nbytes, None)
        for direction, fn in (("fwd", lib.gx_fft_forward), ("inv", lib.gx_fft_inverse)):
            for _ in range(2):
                fn(p, n, dt, None)
            lib.gx_synchronize(None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn(p, n, dt, None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print("n=%d %s %s: %.3f ms per 3-D FFT, %.1f GB/s algorithmic (6B)" % (
                n, np.dtype(dtype).name, direction, ms, 6 * nbytes / ms / 1e6), flush=True)
        lib.gx_free(p)


if __name__ == "__main__":
    main()

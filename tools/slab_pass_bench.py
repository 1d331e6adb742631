#!/usr/bin/env python
"""Times the x / y / z pass kinds on a thin slab at a given line length (one GPU): GB/s of algorithmic traffic (2 B per pass).
    python tools/slab_pass_bench.py 4096 f32 [planes]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genetic_b200 import lib as L
import bench

n = int(sys.argv[1]); dname = sys.argv[2]; P = int(sys.argv[3]) if len(sys.argv) > 3 else 32
lib = L.load(); lib.gx_set_device(0)
dt = L.F32 if dname == "f32" else L.F64
rdt = torch.float32 if dname == "f32" else torch.float64
nc, world = n // 2 + 1, n // P
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
x = torch.randn(P, n, nc, 2, dtype=rdt, device="cuda")
y = torch.empty_like(x)
w = torch.randn(n, P, nc, 2, dtype=rdt, device="cuda")
nbytes = x.numel() * x.element_size()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ty = timed(lambda: lib.gx_slab_fft_y(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, dt, world, 1, 0, 0, st))
tx = timed(lambda: lib.gx_slab_fft_x(C.c_void_p(w.data_ptr()), n, dt, world, 1, 0, st))
tz = timed(lambda: lib.gx_slab_fft_z_c2r(C.c_void_p(y.data_ptr()), n, dt, world, st))
print("n=%d %s planes=%d (%.2f GB): y-pass %.3f ms %.0f GB/s | x-pass %.3f ms %.0f GB/s | z c2r %.3f ms %.0f GB/s | crosscheck %s" % (
    n, dname, P, nbytes / 1e9, ty, 2 * nbytes / ty / 1e6, tx, 2 * nbytes / tx / 1e6, tz, 2 * nbytes / tz / 1e6,
    bench.cufft_crosscheck(lib, torch, n, dt)["ok"]))

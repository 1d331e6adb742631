"""Sweep of column-pass configurations (cluster size x tile width x residency) on the GPU; development aid.
Prints achieved algorithmic GB/s (2 x field bytes per pass) for the x-pass (pass 0) and y-pass (pass 1)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genetic_b200 import lib as L

NAMES = {2048: ["CL8 TC8 minb4 1buf", "CL8 TC8 minb5 1buf", "CL8 TC8 minb6 1buf", "CL16 TC8 minb8 1buf", "CL16 TC8 minb10 1buf", "CL16 TC16 minb4 1buf",
                "CL8 TC8 8pt minb2", "CL8 TC8 8pt minb3", "CL8 TC8 8pt minb4", "CL8 TC8 powers minb5", "CL8 TC8 powers minb4",
                "CL8 TC8 8pt+powers minb3", "CL8 TC8 8pt+powers minb4"],
         1024: ["CL4 TC8 minb4 1buf", "CL4 TC8 minb5 1buf", "CL4 TC8 minb6 1buf", "CL8 TC8 minb8 1buf", "CL8 TC8 minb10 1buf", "CL8 TC16 minb4 1buf"]}


def main():
    import torch
    lib = L.load()
    lib.gx_set_device(0)
    fn = lib.cdll.gx_dev_fft_cols_variant
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    only = None
    args = []
    for a in sys.argv[1:]:
        if a.startswith("--variants="):
            only = [int(x) for x in a.split("=")[1].split(",")]
        else:
            args.append(a)
    for n in [int(a) for a in args] or [1024, 2048]:
        nbytes = lib.gx_field_bytes(n, L.F64)
        p = C.c_void_p()
        lib.gx_malloc(C.byref(p), nbytes)
        lib.gx_memset_zero(p, nbytes, None)
        for v, name in enumerate(NAMES[n]):
            if only is not None and v not in only:
                continue
            for ps, pname in ((0, "x"), (1, "y")):
                rc = fn(p, n, ps, v, None)
                if rc != 0:
                    print("n=%d %-30s pass %s: FAILED %s" % (n, name, pname, lib.gx_last_error().decode()), flush=True)
                    continue
                try:
                    lib.gx_synchronize(None)
                except Exception as e:
                    print("n=%d %-30s pass %s: FAILED %s" % (n, name, pname, e), flush=True)
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    fn(p, n, ps, v, None)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print("n=%d %-30s pass %s: %8.3f ms  %7.1f GB/s" % (n, name, pname, ms, 2 * nbytes / ms / 1e6), flush=True)
        lib.gx_free(p)


if __name__ == "__main__":
    main()

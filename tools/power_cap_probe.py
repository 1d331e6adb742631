"""Does a pass slow down as the board reaches its power cap? The inverse 3-D transform at 2048^3 f64 back to back for a few
seconds after an idle period, with the library's per-class CUDA-event timing and NVML samples of the SM clock and the board
power taken from a side thread. Development tool (the clocks line of bench.py is the contract's sample)."""
import ctypes as C
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genetic_b200 import lib as L


def main():
    import pynvml
    lib = L.load()
    lib.gx_set_device(0)
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n, dt = int(os.environ.get("GX_PROBE_N", "2048")), L.F64
    nbytes = lib.gx_field_bytes(n, dt)
    p = C.c_void_p()
    lib.gx_malloc(C.byref(p), nbytes)
    lib.gx_memset_zero(p, nbytes, None)
    lib.gx_fft_inverse(p, n, dt, None)
    lib.gx_synchronize(None)
    samples, stop = [], [False]

    def sampler():
        while not stop[0]:
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.02)
    time.sleep(3.0)                                   # idle: the board cools down to its boost state
    th = threading.Thread(target=sampler)
    th.start()
    buf = C.create_string_buffer(1 << 14)
    t_start = time.perf_counter()
    for rep in range(int(os.environ.get("GX_PROBE_REPS", "24"))):
        lib.gx_timing_enable(1)
        t0 = time.perf_counter()
        lib.gx_fft_inverse(p, n, dt, None)
        lib.gx_timing_report(buf, len(buf))           # synchronises
        t1 = time.perf_counter()
        cls = {a: float(c) for a, b, c in (ln.split() for ln in buf.value.decode().splitlines())}
        mine = [s for s in samples if t0 <= s[0] <= t1]
        sm = sorted(s[1] for s in mine)[len(mine) // 2] if mine else -1
        pw = max((s[3] for s in mine), default=-1)
        print("t=%.2fs  x %.2f ms  y %.2f ms  z %.2f ms   sm %d MHz  power %.0f W" % (
            t0 - t_start, cls.get("fft_cols_x", 0), cls.get("fft_cols_y", 0), cls.get("fft_rows_c2r", 0), sm, pw), flush=True)
    stop[0] = True
    th.join()
    lib.gx_timing_enable(0)
    lib.gx_free(p)


if __name__ == "__main__":
    main()

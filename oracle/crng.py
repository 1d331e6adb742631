"""TEST INFRASTRUCTURE -- ctypes view of oracle/gx_oracle.c (GSL ranlxs2 + gaussian ziggurat restated).

Follows /root/reference/genetIC/src/simulation/field/randomfieldgenerator.hpp (whole file)."""
import ctypes
import numpy as np
from . import lib


class Stream:
    """One gsl_rng_ranlxs2 stream (randomfieldgenerator.hpp:43-45)."""

    def __init__(self, seed):
        L = lib()
        self._buf = ctypes.create_string_buffer(L.orc_state_size())
        self.seed = int(seed)
        L.orc_ranlxs2_set(self._buf, ctypes.c_ulong(self.seed))

    def gaussians(self, n):
        out = np.empty(int(n), dtype=np.float64)
        lib().orc_gaussians(self._buf, ctypes.c_long(int(n)), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def uniform(self):
        return lib().orc_ranlxs_get_double(self._buf)


def raw_stream(seed, n):
    out = np.empty(int(n), dtype=np.uint32)
    lib().orc_raw_stream(ctypes.c_ulong(int(seed)), ctypes.c_long(int(n)), out.ctypes.data_as(ctypes.c_void_p))
    return out


def draw_real(stream, n, is_float=False):
    """N^3 gaussians in cell order (randomfieldgenerator.hpp:181-192). Returns float64 array [n,n,n]."""
    out = np.empty(n ** 3, dtype=np.float64)
    lib().orc_draw_real(stream._buf, ctypes.c_long(n ** 3), int(is_float), out.ctypes.data_as(ctypes.c_void_p))
    return out.reshape(n, n, n)


def draw_fourier(stream, n, parallel, reverse=False, is_float=False):
    """Square-shell Fourier draw (randomfieldgenerator.hpp:203-260).
    Returns complex128 array [n, n, n//2+1] in the reference's storage order."""
    out = np.empty(2 * n * n * (n // 2 + 1), dtype=np.float64)
    lib().orc_draw_fourier(stream._buf, ctypes.c_ulong(stream.seed), int(n), int(bool(parallel)), int(bool(reverse)),
                           int(is_float), out.ctypes.data_as(ctypes.c_void_p))
    return out.view(np.complex128).reshape(n, n, n // 2 + 1)


def draw_fourier_parallel_shells(seed, n, ks0, ks1, out=None, is_float=False):
    """Timing helper: shells [ks0,ks1) of the parallel draw into a (possibly shared) buffer."""
    if out is None:
        out = np.zeros(2 * n * n * (n // 2 + 1), dtype=np.float64)
    lib().orc_draw_fourier_parallel_shells(ctypes.c_ulong(int(seed)), int(n), int(ks0), int(ks1), int(is_float),
                                           out.ctypes.data_as(ctypes.c_void_p))
    return out

"""TEST INFRASTRUCTURE -- CPU oracle for the genetIC hot path (not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product (genetic_b200) never does; it fails loudly without its CUDA library.

  oracle.crng   -- ctypes binding of gx_oracle.c (ranlxs2 + ziggurat + draw orders)
  oracle.model  -- numpy restatement of the field pipeline (layout, FFT convention, CAMB, Zeldovich, ...)

Parity status: PINNED against the reference's own goldens (tests/test_oracle_golden.py):
test_02k chi^2 text (10 digits), test_13 + test_05a float32 particles bit-exact.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/gx_oracle.c -> oracle/liboracle.so (system gcc; OpenMP if available)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gx_oracle.c")
    hdr = os.path.join(_HERE, "cosmo_impl.inc")
    if (not force and os.path.exists(so)
            and os.path.getmtime(so) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return so
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    base = [cc, "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", so, src, "-lm"]
    try:
        subprocess.check_call(base[:1] + ["-fopenmp"] + base[1:], cwd=_HERE,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except subprocess.CalledProcessError:
        subprocess.check_call(base, cwd=_HERE)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_state_size.restype = ctypes.c_size_t
        _LIB.orc_gaussian_ziggurat.restype = ctypes.c_double
        _LIB.orc_ranlxs_get_double.restype = ctypes.c_double
    return _LIB

"""TEST INFRASTRUCTURE -- CPU oracle for the genetIC hot path (not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product (genetic_b200) never does; it fails loudly without its CUDA library.

  oracle.crng   -- ctypes binding of gx_oracle.c (ranlxs2 + ziggurat + draw orders)
  oracle.model  -- numpy restatement of the field pipeline (layout, FFT convention, CAMB, Zeldovich, ...)

  oracle.multilevel -- numpy restatement of the zoom-level machinery, modifications, splicing
  oracle/_ref/genetIC -- the UNMODIFIED reference compiled from /root/reference against oracle/shims/ (build_ref())

Parity status: PINNED against the reference's own goldens (tests/test_oracle_golden.py, test_oracle_multilevel.py)
and against the reference itself run here on new inputs (tests/test_oracle_ref.py).
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


REF_ROOT = "/root/reference"
REF_BINARY = os.path.join(_HERE, "_ref", "genetIC")


def build_ref(force=False):
    """Compile the UNMODIFIED reference from /root/reference against oracle/shims/ (FFTW3 / GSL / HDF5 stand-ins, see
    oracle/shims/README.md) into oracle/_ref/genetIC. Only possible where /root/reference exists (the build container);
    elsewhere the prebuilt binary is used if it travelled with the snapshot. Returns the path or None."""
    if os.path.isdir(os.path.join(REF_ROOT, "genetIC", "src")):
        srcs = [os.path.join(_HERE, "build_ref.sh"), os.path.join(_HERE, "gx_oracle.c")] + \
               [os.path.join(_HERE, "shims", f) for f in ("fft_shim.cpp", "gsl_shim.cpp", "fftw3.h", "highfive/highfive.hpp")]
        newest = max(os.path.getmtime(f) for f in srcs)
        have_cufft = os.path.exists("/usr/local/cuda/include/cufftw.h")
        stale = not os.path.exists(REF_BINARY) or os.path.getmtime(REF_BINARY) < newest or \
            (have_cufft and (not os.path.exists(REF_BINARY + "_cufft") or os.path.getmtime(REF_BINARY + "_cufft") < newest))
        if force or stale:
            subprocess.check_call(["/bin/sh", os.path.join(_HERE, "build_ref.sh"), REF_ROOT], stdout=subprocess.DEVNULL)
    return ref_binary()


def ref_binary(cufft=False):
    """Path of the compiled reference (oracle/_ref/genetIC; cufft=True: oracle/_ref/genetIC_cufft, the reference's own
    USE_CUFFT=1 GPU path through the toolkit's cuFFTW) if it exists and is executable, else None."""
    exe = REF_BINARY + ("_cufft" if cufft else "")
    return exe if os.access(exe, os.X_OK) else None


def run_ref(paramfile_text, workdir, files=(), timeout=3600, args=(), cufft=False, threads=None):
    """Runs the compiled reference on a paramfile written into `workdir` (extra input files are copied next to it).
    Returns its combined stdout/stderr. Raises if the binary is missing or exits non-zero."""
    import shutil
    exe = ref_binary(cufft)
    if exe is None:
        raise RuntimeError("oracle/_ref/genetIC has not been built (needs /root/reference: oracle.build_ref())")
    os.makedirs(workdir, exist_ok=True)
    for f in files:
        shutil.copy(f, os.path.join(workdir, os.path.basename(f)))
    with open(os.path.join(workdir, "paramfile.txt"), "w") as fh:
        fh.write(paramfile_text)
    env = dict(os.environ)
    if threads is not None:  # torchrun exports OMP_NUM_THREADS=1: the reference's thread count is set, not inherited
        env["OMP_NUM_THREADS"] = str(int(threads))
    r = subprocess.run([exe, "paramfile.txt"] + list(args), cwd=workdir, capture_output=True, text=True, timeout=timeout,
                       env=env)
    if r.returncode != 0:
        raise RuntimeError("reference binary failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return r.stdout + r.stderr


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_state_size.restype = ctypes.c_size_t
        _LIB.orc_gaussian_ziggurat.restype = ctypes.c_double
        _LIB.orc_ranlxs_get_double.restype = ctypes.c_double
    return _LIB

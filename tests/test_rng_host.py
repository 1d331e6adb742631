"""CPU tests of the device RNG core (genetic_b200/csrc/rng_core.cuh compiled for the host): skip-ahead algebra,
raw stream and ziggurat parse against the oracle's sequential GSL restatement. No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import crng

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "rng_core_host.cu")
    so = os.path.join(HERE, "host", "_rng_core_host.so")
    hdr = os.path.join(HERE, "..", "genetic_b200", "csrc", "rng_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-ccbin",
                               "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, src],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ctypes.CDLL(so)


def test_lcg_equivalence_and_jump(hostlib):
    for seed in (1, 42, 8896131, 2 ** 31 - 1):
        assert hostlib.host_rng_selfcheck(ctypes.c_ulong(seed)) == 0


@pytest.mark.parametrize("seed,offset", [(42, 0), (42, 1), (42, 23), (42, 24), (8896131, 24 * 1000 + 5),
                                         (13842314, 123456789), (7, 24 * 12 * 397)])
def test_raw_stream_with_skip_ahead(hostlib, seed, offset):
    n = 500
    out = np.empty(n, dtype=np.uint32)
    hostlib.host_rng_raw(ctypes.c_ulong(seed), ctypes.c_ulonglong(offset), ctypes.c_long(n),
                         out.ctypes.data_as(ctypes.c_void_p))
    if offset < 3_000_000:
        ref = crng.raw_stream(seed, offset + n)[offset:]
    else:
        ref = crng.raw_stream(seed, offset + n)[offset:]
    assert np.array_equal(out, ref)


def test_ziggurat_parse_matches_gsl(hostlib):
    seed, ng = 13842314, 200000
    raw = crng.raw_stream(seed, int(ng * 2.1) + 100)
    out = np.empty(ng)
    hostlib.host_rng_gaussians_from_raw.restype = ctypes.c_long
    used = hostlib.host_rng_gaussians_from_raw(raw.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(raw.size),
                                               ctypes.c_long(ng), out.ctypes.data_as(ctypes.c_void_p))
    assert used > 0
    ref = crng.Stream(seed).gaussians(ng)
    assert np.array_equal(out, ref)  # host build: same libm, so even the tail values agree bit for bit


@pytest.fixture(scope="module")
def stagedlib():
    src = os.path.join(HERE, "host", "rng_staged_host.cu")
    so = os.path.join(HERE, "host", "_rng_staged_host.so")
    hdrs = [os.path.join(HERE, "..", "genetic_b200", "csrc", h) for h in ("rng_core.cuh", "rng_staged_core.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-ccbin",
                               "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, src],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("n,seed,reverse", [(8, 42, 0), (16, 8896131, 0), (32, 8896131, 1), (64, 1512351, 0)])
def test_staged_noise_path_on_the_host(stagedlib, n, seed, reverse):
    """The staged single-generation noise path (rng_staged_core.cuh: hypothesis lists, resolve, call-driven delivery with
    the analytic overwrite rule, never-drawn set) run on the CPU: bit-exact against the oracle's sequential draw."""
    nc = n // 2 + 1
    out = np.zeros((n, n, nc), dtype=np.complex128)
    written = np.zeros((n, n, nc), dtype=np.uint8)
    stats = np.zeros(8, dtype=np.int64)
    rc = stagedlib.host_staged_noise(n, ctypes.c_ulong(seed), reverse, out.ctypes.data_as(ctypes.c_void_p),
                                     written.ctypes.data_as(ctypes.c_void_p), stats.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    ref = crng.draw_fourier(crng.Stream(seed), n, True, bool(reverse), False)   # parallel order, optionally reversed draws
    assert np.array_equal(out, ref)

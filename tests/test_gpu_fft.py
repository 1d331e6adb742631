"""GPU parity: the hand-written 3-D FFT and the spectral operators against the CPU oracle (oracle/model.py).
All calls go through the C ABI (genetic_b200.lib). Tolerances: f64 1e-10 relative (north star), f32 1e-5."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import genetic_b200 as gb
from oracle import model as M

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tol(dtype):
    return 1e-10 if np.dtype(dtype) == np.float64 else 1e-5


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256])
def test_fft_forward_inverse_vs_oracle(n, dtype):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, n, n)).astype(dtype)
    g = gb.Grid(100.0, n, dtype=dtype)
    f = gb.Field(g, x, fourier=False)
    f.toFourier()
    F = f.getDataVector()
    Fo = M.to_fourier(x.astype(np.float64))
    assert _rel(F, Fo) < _tol(dtype)
    # the transform is unitary: Parseval through the reference's mode weights
    assert abs(np.sum(M.mode_weights(n) * np.abs(F) ** 2) / np.sum(x.astype(np.float64) ** 2) - 1) < 10 * _tol(dtype)
    f.toReal()
    y = f.getDataVector()
    assert _rel(y, x) < _tol(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_inverse_of_unmirrored_spectrum(dtype):
    """toReal mirrors first (fourier.hpp:497): fill only independent slots, redundant ones hold garbage."""
    n = 16
    rng = np.random.default_rng(5)
    ct = np.complex128 if dtype == np.float64 else np.complex64
    F = (rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))).astype(ct)
    g = gb.Grid(100.0, n, dtype=dtype)
    f = gb.Field(g, F, fourier=True)
    f.ensureFourierModesAreMirrored()
    Fm = f.getDataVector()
    assert np.array_equal(Fm, M.mirror(F.copy()))  # bit-exact storage, including the self-conjugate sign flips
    f2 = gb.Field(g, F, fourier=True)
    f2.toReal()
    assert _rel(f2.getDataVector(), M.to_real(F.astype(np.complex128))) < _tol(dtype)


@pytest.mark.parametrize("n", [128, 512])
def test_fft_cross_check_against_cufft(n):
    """In-run cross-check of the hand-written passes against cuFFT (through torch.fft on the same device): the north star
    allows cuFFT for exactly this. Forward r2c and inverse c2r, f64, unitary scaling n^1.5 (fourier.hpp:490,512-513)."""
    import torch
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal((n, n, n))
    g = gb.Grid(100.0, n)
    f = gb.Field(g, x, fourier=False)
    f.toFourier()
    F = f.getDataVector()
    xt = torch.from_numpy(x).cuda()
    Ft = (torch.fft.rfftn(xt) / float(n) ** 1.5).cpu().numpy()
    assert _rel(F, Ft) < 1e-13
    f.toReal()
    yt = (torch.fft.irfftn(torch.from_numpy(Ft).cuda(), s=(n, n, n)) * float(n) ** 1.5).cpu().numpy()
    assert _rel(f.getDataVector(), yt) < 1e-13


@pytest.mark.parametrize("n", [512])
def test_fft_roundtrip_large(n):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, n, n))
    g = gb.Grid(100.0, n)
    f = gb.Field(g, x, fourier=False)
    f.toFourier()
    F = f.getDataVector()
    # spot-check a few modes against a direct DFT sum (size-independent property: linearity + known modes)
    for (kx, ky, kz) in [(0, 0, 0), (1, 2, 3), (n // 2, 0, 5), (n - 1, n - 2, n // 2)]:
        ph = np.exp(-2j * np.pi * (kx * np.arange(n)[:, None, None] + ky * np.arange(n)[None, :, None]
                                   + kz * np.arange(n)[None, None, :]) / n)
        ref = (x * ph).sum() / n ** 1.5
        assert abs(F[kx, ky, kz] - ref) < 1e-10 * np.sqrt(n ** 3)
    f.toReal()
    assert _rel(f.getDataVector(), x) < 1e-12


@pytest.mark.parametrize("n", [512, 1024])
def test_fft_cluster_passes_are_race_free(n):
    """The column passes of n > 256 exchange data between the CTAs of a thread-block cluster through distributed
    shared memory; a missing release/acquire there shows up as sparse wrong tiles that differ from run to run.
    Bit-identical repeats + Parseval + the inverse returning the input pin it at sizes numpy cannot check quickly."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, n, n))
    g = gb.Grid(100.0, n)
    outs = []
    for rep in range(3):
        f = gb.Field(g, x, fourier=False)
        f.toFourier()
        outs.append(f.getDataVector())
        if rep == 2:
            f.toReal()
            assert _rel(f.getDataVector(), x) < 1e-12
        del f
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    w = np.full(n // 2 + 1, 2.0)
    w[0] = w[-1] = 1.0
    assert abs(np.sum(np.abs(outs[0]) ** 2 * w) / np.sum(x ** 2) - 1) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_transfer_zeldovich_emit_vs_oracle(dtype):
    """White noise from the oracle (device noise is tested separately) -> sqrt(P) -> psi_xyz -> particles."""
    n, box = 32, 50.0
    cos_o = M.Cosmology(0.279, 0.721, 0.817, 99)
    camb_file = os.path.join(GOLD, "camb_transfer_kmax40_z0.dat")
    r = M.unigrid_ic(n, box, 8896131, "fourier_parallel", camb_file, cos_o, dtype=dtype, return_intermediate=True)
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    g = gb.Grid(box, n, dtype=dtype)
    spec = gb.CAMB(cos, camb_file, "all", dtype)
    assert abs(spec.amplitude / r["camb"].amplitude - 1) < (1e-13 if dtype == np.float64 else 1e-5)
    cov = spec.getPowerSpectrumForGrid(g).getDataVector()
    # (the spline, exp, log and pow behind C(k) are the device's libm against the host's: measured agreement ~1e-13 in double)
    assert _rel(cov.real, r["cov"]) < (_tol(dtype) if dtype == np.float64 else 10 * _tol(dtype)) and np.all(cov.imag == 0)
    f = gb.Field(g, r["white"], fourier=True)
    f.applyTransferFunction(spec, 0.5)
    D = f.getDataVector()
    assert _rel(D, r["delta_k"]) < _tol(dtype) * 10
    # the Zeldovich operator itself is pure IEEE arithmetic: bit-exact given identical delta_k
    f2 = gb.Field(g, r["delta_k"], fourier=True)
    for axis in range(3):
        psi = gb.Field(g, fourier=True)
        f2.lib.gx_zeldovich_fourier(f2.ptr, psi.ptr, n, f2.dt, float(g.thisGridSize), axis, None)
        assert np.array_equal(psi.getDataVector(), r["psi_k"][axis])
    gen = gb.ZeldovichParticleGenerator(f2)
    for axis in range(3):
        assert _rel(gen.getGeneratedFields()[axis].getDataVector(), r["psi"][axis]) < _tol(dtype) * 10
    from genetic_b200.host import emit_particles
    pos, vel = emit_particles(gen.getGeneratedFields(), g, cos, out_dtype=dtype)
    assert np.abs(vel - r["vel"]).max() < _tol(dtype) * 10 * np.abs(r["vel"]).max()
    d = np.abs(pos - r["pos"])
    d = np.minimum(d, box - d)  # a particle within rounding of the box edge may wrap differently
    assert d.max() < _tol(dtype) * 10 * box


def test_reductions_and_axpy():
    n = 32
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
    B = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
    g = gb.Grid(10.0, n)
    fa, fb = gb.Field(g, A), gb.Field(g, B)
    ref = np.sum(M.mode_weights(n) * (np.conj(A) * B).real)
    assert abs(fa.innerProduct(fb) - ref) < 1e-12 * abs(ref) + 1e-9
    fa.addScaledFourier(fb, 0.25)
    ind = M.independent_mask(n)
    out = fa.getDataVector()
    assert np.allclose(out[ind], (A + 0.25 * B)[ind], rtol=1e-15, atol=0) and np.array_equal(out[~ind], A[~ind])
    x = rng.standard_normal((n, n, n))
    y = rng.standard_normal((n, n, n))
    fx, fy = gb.Field(g, x, fourier=False), gb.Field(g, y, fourier=False)
    assert abs(fx.innerProduct(fy) - np.sum(x * y)) < 1e-10
    fx *= 3.0
    assert np.array_equal(fx.getDataVector(), x * 3.0)


def test_filters_and_unit_variance():
    n, box = 16, 20.0
    rng = np.random.default_rng(9)
    A = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
    g = gb.Grid(box, n)
    kmin = 2 * np.pi / box
    ks = M.signed_modes(n) * kmin
    k = np.sqrt(ks[:, None, None] ** 2 + ks[None, :, None] ** 2 + (np.arange(n // 2 + 1) * kmin)[None, None, :] ** 2)
    kcut = 0.3 * n * kmin
    lp = 1.0 / (1.0 + np.exp((k - kcut) / (0.1 * kcut)))
    ind = M.independent_mask(n)
    f = gb.Field(g, A)
    f.applyFilter(gb.LowPassFermiFilter(kcut))
    assert np.allclose(f.getDataVector()[ind], (A * lp)[ind], rtol=1e-13)
    from genetic_b200.host import ComplementaryCovarianceFilter
    f = gb.Field(g, A)
    f.applyFilter(ComplementaryCovarianceFilter(kcut))
    assert np.allclose(f.getDataVector()[ind], (A * np.sqrt(1 - lp * lp))[ind], rtol=1e-12, atol=1e-15)
    f = gb.Field(g, A)
    f.enforceUnitVariance()
    assert np.allclose(np.abs(f.getDataVector()[ind]), 1.0, rtol=1e-14)
    # against the oracle's statement of MultiLevelField::enforceUnitVariance (multilevelfield.hpp:498-510): delta_k / |delta_k|
    U = f.getDataVector()
    ref = A / np.abs(A)
    assert np.abs(U[ind] - ref[ind]).max() <= 4 * np.finfo(np.float64).eps
    Z = A.copy()
    Z[3, 5, 2] = 0.0                                    # 0 stays 0
    f = gb.Field(g, Z)
    f.enforceUnitVariance()
    assert f.getDataVector()[3, 5, 2] == 0


def test_errors_are_raised():
    g = gb.Grid(10.0, 9)                                # odd sizes are refused (even sizes: powers of two or mixed radix)
    with pytest.raises(gb.GxError):
        f = gb.Field(g, fourier=False)
        f.toFourier()


@pytest.mark.parametrize("n,dtype", [(16, np.float64), (32, np.float64), (32, np.float32)])
def test_power_spectrum_bins_on_device(n, dtype):
    """gx_power_spectrum_bins (the reduction behind `dump_ps`, cosmology/parameters.hpp:112-205) against the reference's
    triple loop over [-n/2, n/2]^3 restated in numpy (tests/host/mock_device.py): counts exact, sums to rounding."""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location("gx_mock_device", os.path.join(os.path.dirname(__file__), "host", "mock_device.py"))
    mock = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mock)
    box = 37.0
    rng = np.random.default_rng(n)
    ct = np.complex128 if dtype == np.float64 else np.complex64
    A = (rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))).astype(ct)
    P = (rng.random((n, n, n // 2 + 1)) + 0.5).astype(ct)
    g = gb.Grid(box, n, dtype=dtype)
    f, cov = gb.Field(g, A, fourier=True), gb.Field(g, P, fourier=True)
    bins = np.zeros((4, 100))
    f.lib.gx_power_spectrum_bins(f.ptr, cov.ptr, n, f.dt, box, 100, bins.ctypes.data_as(C.c_void_p), None)
    m = mock.MockLib()
    hp, hc = C.c_void_p(), C.c_void_p()
    m.gx_malloc(C.byref(hp), m.gx_field_bytes(n, 1))
    m.gx_malloc(C.byref(hc), m.gx_field_bytes(n, 1))
    m.F(hp, n)[:] = A.astype(np.complex128)
    m.F(hc, n)[:] = P.astype(np.complex128)
    ref = np.zeros((4, 100))
    m.gx_power_spectrum_bins(hp, hc, n, 1, box, 100, ref.ctypes.data_as(C.c_void_p), None)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    if dtype == np.float64:
        assert np.array_equal(bins[0], ref[0])
    else:   # single precision: a mode within rounding of a bin edge may fall on the other side
        assert np.abs(bins[0] - ref[0]).sum() <= 8 and bins[0].sum() == ref[0].sum()
        return
    for i in (1, 2, 3):
        assert np.allclose(bins[i], ref[i], rtol=tol, atol=0)


# ---- grid sizes that are not powers of two: the mixed-radix passes (csrc/fft_generic.cu) ---------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [6, 12, 20, 24, 48, 60, 66, 96, 100, 144])
def test_mixed_radix_fft_vs_oracle(n, dtype):
    """The reference takes any N through FFTW (fourier.hpp:521-563): radices 4, 2, 3, 5 and larger primes (66 = 2 * 3 * 11)."""
    test_fft_forward_inverse_vs_oracle(n, dtype)


@pytest.mark.parametrize("n", [12, 36])
def test_mixed_radix_inverse_mirrors_first(n):
    rng = np.random.default_rng(n)
    F = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
    g = gb.Grid(100.0, n)
    f = gb.Field(g, F, fourier=True)
    f.toReal()
    assert _rel(f.getDataVector(), M.to_real(F)) < 1e-10


def test_unigrid_48_real_space_seed_vs_oracle():
    """Whole chain at 48^3 (real-space draw order, transfer, Zeldovich operator as a sweep + the mixed-radix inverse, particle
    evaluation) against the oracle on the same seed."""
    n, box, seed = 48, 100.0, 8896131
    camb = os.path.join(GOLD, "planck_2018_transfer_out.dat")
    cos = gb.CosmologicalParameters(0.3158, 0.6842, 0.8117, 200, ns=0.9660, hubble=0.6732)
    pos, vel, mass, eps = gb.generate_unigrid_ic(n, box, seed, cos, camb, mode="real", out_dtype=np.float64)
    ocos = M.Cosmology(0.3158, 0.6842, 0.8117, 200, ns=0.9660, hubble=0.6732)
    opos, ovel, omass, oeps = M.unigrid_ic(n, box, seed, "real", camb, ocos)
    assert np.abs(vel - ovel).max() < 1e-10 * np.abs(ovel).max()
    d = np.abs(pos - opos)
    d = np.minimum(d, box - d)
    assert d.max() < 1e-10 * box
    assert mass == omass and eps == oeps


def test_fourier_draw_orders_refuse_mixed_radix_sizes():
    g = gb.Grid(100.0, 48)
    f = gb.Field(g, fourier=True)
    rfg = gb.RandomFieldGenerator(f)
    rfg.seed(1)
    rfg.setDrawInFourierSpace(True)
    with pytest.raises(gb.GxError, match="power-of-two"):
        rfg.draw()

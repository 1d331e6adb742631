"""Pins the CPU oracle (oracle/) against the reference's own golden files.

CPU-only. The fixtures under tests/golden/ are verbatim data files from the reference's test suite
(see tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import crng, model as M


def _cosmo_279():
    return M.Cosmology(0.279, 0.721, 0.817, 99)


def test_gsl_known_answers():
    # first gaussians of gsl_ran_gaussian_ziggurat(ranlxs2) for seeds 42 and 43
    g = crng.Stream(42).gaussians(4)
    assert g.tolist() == [1.6658629784446979, 0.7826561053875941, -0.9834597415502253, -0.06695916569209004]
    g = crng.Stream(43).gaussians(4)
    assert g.tolist() == [0.925654650073142, 1.4551845794392302, 1.330013772424191, 0.5866959993899714]
    # gsl_rng_set maps seed 0 -> 1
    assert np.array_equal(crng.raw_stream(0, 100), crng.raw_stream(1, 100))
    # seeds are masked to 31 bits
    assert np.array_equal(crng.raw_stream(5, 100), crng.raw_stream(5 + 2 ** 31, 100))
    r = crng.raw_stream(42, 10000)
    assert r.max() < 2 ** 24


def test_chi2_golden_test_02k(golden_dir):
    # tests/test_02k_chisquare_onelevel: random_seed_real_space 13842314, 64^3
    ref = open(os.path.join(golden_dir, "test_02k_chisquare_onelevel", "reference.txt")).read()
    W = M.draw_white_noise(64, 13842314, "real")
    chi2 = M.chi2_unigrid(W)
    line = "Calculated chi^2 = %.10g (dof = %d)" % (chi2, 64 ** 3)
    assert line == ref.strip()


@pytest.mark.parametrize("test,box,seed,mode", [
    ("test_13_parallel_seed", 10.0, 42, "fourier_parallel"),
    ("test_05a_seed_old", 50.0, 8896131, "fourier_reverse"),
])
def test_tipsy_goldens_bit_exact(golden_dir, test, box, seed, mode):
    cos = _cosmo_279()
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    pos, vel, mass, eps = M.unigrid_ic(8, box, seed, mode, camb, cos)
    t = M.to_tipsy_units(pos, vel, mass, eps, box, cos, 8 ** 3)
    g = M.read_tipsy(os.path.join(golden_dir, test, "reference_output"))
    assert np.array_equal(t["pos"], g["pos"])
    assert np.array_equal(t["vel"], g["vel"])
    assert np.all(g["mass"] == t["mass"]) and np.all(g["eps"] == t["eps"])


def test_test13_intermediate_known_answers(golden_dir):
    # SURVEY.md 8(c) "test_13 intermediate known answers"
    cos = _cosmo_279()
    c = M.CAMB(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), cos)
    assert M.growth_factor(0.279, 0.721, 0.01) == 0.0099999946843276344
    assert M.growth_factor(0.279, 0.721, 1.0) == 0.76445955115974173
    # (value from the survey's numpy/scipy probe: different libm/spline code path, agree to ~1e-14)
    assert abs(c.amplitude / 645.54324716013275 - 1) < 1e-12
    assert abs(c.transfer(1.0) / 0.0041117270188788914 - 1) < 1e-12
    assert abs(M.velocity_to_offset_ratio(cos) / 5282.0518835013354 - 1) < 1e-15


def test_older_golden_envelopes(golden_dir):
    """test_01a (64^3 real-space seed, float32 tipsy) agrees to <= 2 float32 ulp; test_19 grid to 5e-6 sigma.
    These older goldens predate a change in the reference's transfer-function code."""
    cos = _cosmo_279()
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    pos, vel, mass, eps = M.unigrid_ic(64, 50.0, 8896131, "real", camb, cos)
    t = M.to_tipsy_units(pos, vel, mass, eps, 50.0, cos, 64 ** 3)
    z = np.load(os.path.join(golden_dir, "test_01a_basic_grid", "particles_every61.npz"))
    idx, g = z["index"], z["particles"]
    assert np.abs(t["pos"][idx] - g[:, 1:4]).max() < 3e-7
    assert np.abs(t["vel"][idx] - g[:, 4:7]).max() < 3e-5 * np.abs(g[:, 4:7]).max()
    assert (t["pos"][idx] == g[:, 1:4]).mean() > 0.95

    cos19 = M.Cosmology(0.3156, 0.6844, 0.831, 500, ns=0.9645, hubble=0.6727)
    r = M.unigrid_ic(64, 64.0, 8896131, "real", os.path.join(golden_dir, "camb_transfer_kmax40_z500.dat"), cos19,
                     return_intermediate=True)
    delta = M.to_real(r["delta_k"])
    gg = np.load(os.path.join(golden_dir, "test_19_rounding_errors", "grid0_every4.npz"))["grid"]
    assert np.abs(delta[::4, ::4, ::4] - gg).max() < 1e-5 * gg.std()


def test_mirror_fast_matches_sequential():
    rng = np.random.default_rng(0)
    for n in (4, 8):
        F = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
        assert np.array_equal(M.mirror(F.copy()), M.mirror_fast(F.copy()))


def test_mode_weights_sum_to_n3():
    # every one of the N^3 modes of the full cube is represented exactly once through the weights
    for n in (4, 8, 16):
        assert M.mode_weights(n).sum() == n ** 3

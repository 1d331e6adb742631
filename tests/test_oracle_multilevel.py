"""Pins the multi-level / modification / splicing oracle (oracle/multilevel.py) against the reference's own goldens
(tests/golden/, copied from /root/reference/genetIC/tests by make_golden.py). CPU only."""
import os
import re

import numpy as np

from oracle import model as M, multilevel as ML


def _cosmo():
    return M.Cosmology(0.279, 0.721, 0.817, 99)


def _camb(golden_dir):
    return M.CAMB(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), _cosmo())


def _ids(golden_dir, test):
    return np.sort(np.loadtxt(os.path.join(golden_dir, test, "zoomtest.txt"), dtype=np.int64).reshape(-1))


def test_chi2_two_levels_golden_02m(golden_dir):
    """test_02m: basegrid 64 64; centre 25 32 24; select_cube 16; zoomgrid 2 64; chi2 -- exercises the zoom window
    placement (ic.hpp:398-479), the windowed covector metric (multilevelfield.hpp:346-391), coarse->fine tricubic
    interpolation and fine->coarse averaging."""
    grids = [ML.Grid(64.0, 64)]
    ML.select_cube(grids, (25, 32, 24), 16)
    assert ML.zoom_corner(grids[0], 2) == (8, 15, 7)
    ML.add_zoom_level(grids, 2, 64)
    fields = ML.draw_levels(grids, 13842314, "real")
    text = open(os.path.join(golden_dir, "test_02m_chisquare_twolevels", "reference.txt")).read()
    ref, dof = re.search(r"chi\^2 = ([0-9.]+) \(dof = (\d+)\)", text).groups()
    assert ML.num_dof(grids) == int(dof)
    assert "%.10g" % ML.chi2(grids, fields) == ref


def test_zoom_combination_golden_01c(golden_dir):
    """test_01c: 64^3 base + zoom x3 64^3, real-space seed; dump_grid 0/1 after `done` = the per-level overdensity after
    initialiseParticleGeneratorBasedOnTemplate (multilevelgenerator.hpp:79-128). Older golden: 5e-7 sigma envelope
    (same origin as test_01a/test_19)."""
    grids = [ML.Grid(50.0, 64)]
    grids[0].flags = _ids(golden_dir, "test_01c_zoom_grid")
    ML.add_zoom_level(grids, 3, 64)
    assert grids[1].offset == (16.40625, 16.40625, 16.40625)
    camb = _camb(golden_dir)
    fields = ML.draw_levels(grids, 8896131, "real")
    delta = [M.apply_transfer(f, camb.covariance(g.size, g.n, "all"), 0.5) for f, g in zip(fields, grids)]
    out = ML.combine_levels(grids, delta)
    z = np.load(os.path.join(golden_dir, "test_01c_zoom_grid", "grids_every4.npz"))
    for l in range(2):
        ref = z["grid%d" % l]
        assert np.abs(out[l][0][::4, ::4, ::4] - ref).max() < 1e-6 * ref.std()


def test_two_linear_modifications_with_zoom_golden_02c(golden_dir):
    """test_02c: two `modify overdensity relative` on single cells of a x3 zoom; grids after `done`. Exercises the
    deinterpolated multi-level covectors (multilevelgrid.hpp:245-321), Gram-Schmidt and the linear update
    (modificationmanager.hpp:199-217,283-304). Current-generation golden: agrees to 2e-9 sigma."""
    camb = _camb(golden_dir)
    grids = [ML.Grid(64.0, 64)]
    grids[0].flags = _ids(golden_dir, "test_02c_two_modifs")
    ML.add_zoom_level(grids, 3, 64)
    f = ML.draw_levels(grids, 8896131, "real")
    cdm = [camb.covariance(g.size, g.n, "dm") for g in grids]
    call = [camb.covariance(g.size, g.n, "all") for g in grids]
    alphas, targets = [], []
    for centre, rel in (((32.5, 31.5, 32.5), 10.0), ((32.5, 33.5, 32.5), -10.0)):
        ML.select_nearest(grids, centre)
        a = ML.overdensity_covector(grids, cdm)
        alphas.append(a)
        targets.append(rel * ML.inner_product(grids, a, f))
    f2, _ = ML.apply_linear_modifications(grids, f, alphas, targets)
    out = ML.combine_levels(grids, [M.apply_transfer(x, c, 0.5) for x, c in zip(f2, call)])
    z = np.load(os.path.join(golden_dir, "test_02c_two_modifs", "grids_every4.npz"))
    for l in range(2):
        ref = z["grid%d" % l]
        assert np.abs(out[l][0][::4, ::4, ::4] - ref).max() < 1e-8 * ref.std()


def test_variance_and_overdensity_values_golden_02f_02g(golden_dir):
    camb = _camb(golden_dir)
    grids = [ML.Grid(64.0, 64)]
    ML.select_cube(grids, (8, 8, 8), 20)
    f = ML.draw_levels(grids, 8896131, "real")
    covs = [camb.covariance(g.size, g.n, "dm") for g in grids]
    ref = open(os.path.join(golden_dir, "test_02f_onelevelvariance", "reference.txt")).read()
    assert "variance: calculated value = %.6g" % ML.variance_value(grids, covs, f, 10000.) in ref

    grids = [ML.Grid(64.0, 64)]
    ML.select_cube(grids, (8, 8, 8), 16)
    ML.add_zoom_level(grids, 4, 64, corner=(0, 0, 0))   # zoomgrid_with_origin_at 0 0 0 4 64
    ML.select_cube(grids, (8, 8, 8), 16)
    f = ML.draw_levels(grids, 8896131, "real")
    covs = [camb.covariance(g.size, g.n, "dm") for g in grids]
    ref = open(os.path.join(golden_dir, "test_02g_multilevelvariance", "reference.txt")).read()
    a = ML.overdensity_covector(grids, covs)
    assert "overdensity: calculated value = %.6g" % ML.inner_product(grids, a, f) in ref
    assert "variance: calculated value = %.6g" % ML.variance_value(grids, covs, f, 10000.) in ref


def test_overdensity_and_variance_modification_golden_02i(golden_dir):
    """test_02i: zoom x4, `modify overdensity relative 1` + `modify variance relative 2` with filtering_scale 2 and the
    default 10 initial steps / 0.1 % precision; grids after `done`. Pins performIterations, addToOrthonormalFamily and
    the step-count rule (modificationmanager.hpp:219-281) end to end."""
    camb = _camb(golden_dir)
    grids = [ML.Grid(64.0, 64)]
    ML.select_sphere(grids, (32.5, 32.5, 32.5), 5)
    ML.add_zoom_level(grids, 4, 64)
    ML.select_cube(grids, (32.5, 32.5, 32.5), 6.0)
    f = ML.draw_levels(grids, 8896131, "real")
    cdm = [camb.covariance(g.size, g.n, "dm") for g in grids]
    call = [camb.covariance(g.size, g.n, "all") for g in grids]
    a = ML.overdensity_covector(grids, cdm)
    v = ML.inner_product(grids, a, f)
    var0 = ML.variance_value(grids, cdm, f, 2.0)
    f, vecs = ML.apply_linear_modifications(grids, f, [a], [1.0 * v])
    f = ML.apply_variance_modification(grids, cdm, f, vecs, 2.0 * var0, 2.0)
    assert abs(ML.variance_value(grids, cdm, f, 2.0) / (2.0 * var0) - 1) < 1e-3
    out = ML.combine_levels(grids, [M.apply_transfer(x, c, 0.5) for x, c in zip(f, call)])
    z = np.load(os.path.join(golden_dir, "test_02i_multilevel_algorithm", "grids_every4.npz"))
    for l in range(2):
        ref = z["grid%d" % l]
        assert np.abs(out[l][0][::4, ::4, ::4] - ref).max() < 1e-6 * ref.std()


def _unigrid_particles(g, f, camb, cos):
    D = M.apply_transfer(f, camb.covariance(g.size, g.n, "all"), 0.5)
    psis = [M.to_real(p) for p in M.zeldovich_offsets_fourier(D, g.size)]
    return M.particles_unigrid(psis, g.size, cos)


def test_velocity_modification_golden_14(golden_dir):
    """test_14: basegrid 50 8; idfile; modify vx absolute 0 -- VelocityModification with the real-space 4th-order
    derivative (linearmodification.hpp:186-285); gadget3 golden (float32)."""
    camb, cos = _camb(golden_dir), _cosmo()
    g = ML.Grid(50.0, 8)
    g.flags = np.sort(np.loadtxt(os.path.join(golden_dir, "test_14_velocity_modif", "ids.txt"), dtype=np.int64))
    f = ML.draw_levels([g], 8896131, "real")
    cdm = [camb.covariance(g.size, g.n, "dm")]
    a = ML.velocity_covector([g], cdm, 0, M.velocity_to_offset_ratio(cos))
    f2, _ = ML.apply_linear_modifications([g], f, [a], [0.0])
    pos, vel, _, _ = _unigrid_particles(g, f2[0], camb, cos)
    z = np.load(os.path.join(golden_dir, "test_14_velocity_modif", "particles.npz"))
    assert np.abs(pos - z["pos"]).max() < 4e-6 and np.abs(vel - z["vel"]).max() < 1e-4


def test_angular_momentum_modification_golden_22(golden_dir):
    """test_22: base_grid 50 32; select_ellipse 10 7.5 5; modify Ly absolute 100 -- AngMomentumModification
    (linearmodification.hpp:288-411), Grid::getCentreWrapped; gadget3 golden (float32)."""
    camb, cos = _camb(golden_dir), _cosmo()
    g = ML.Grid(50.0, 32)
    ML.select_ellipse([g], (25, 25, 25), 10, 7.5, 5)
    f = ML.draw_levels([g], 8896131, "real")
    cdm = [camb.covariance(g.size, g.n, "dm")]
    a = ML.angular_momentum_covector([g], cdm, 1, M.velocity_to_offset_ratio(cos))
    f2, _ = ML.apply_linear_modifications([g], f, [a], [100.0])
    pos, vel, _, _ = _unigrid_particles(g, f2[0], camb, cos)
    z = np.load(os.path.join(golden_dir, "test_22_angular_momentum_modif", "particles_every7.npz"))
    assert np.abs(pos[z["index"]] - z["pos"]).max() < 4e-6 and np.abs(vel[z["index"]] - z["vel"]).max() < 1e-4


def test_splice_golden_24(golden_dir):
    """test_24: 16^3, select_sphere 10, splice 8896132 -- spliceOneLevel + conjugateGradient (splice.hpp:38-118,
    cg.hpp:14-58). CG stops at rtol 1e-6, so the grid is pinned to a few 1e-6 sigma."""
    camb = _camb(golden_dir)
    g = ML.Grid(50.0, 16)
    ML.select_sphere([g], (25, 25, 25), 10)
    assert len(g.flags) == 136
    orig = ML.draw_levels([g], 8896131, "real")[0]
    new = ML.draw_levels([g], 8896132, "real")[0]
    cov = camb.covariance(g.size, g.n, "all")
    sp = ML.splice_one_level(g, new, orig, cov)
    delta = M.to_real(M.apply_transfer(sp, cov, 0.5))
    ref = np.load(os.path.join(golden_dir, "test_24_splicing", "grid0.npz"))["grid0"]
    assert np.abs(delta - ref).max() < 1e-5 * ref.std()
    # inside the flagged region the result reproduces field `a` of spliceOneLevel (the one drawn from the new seed,
    # ic.hpp:1681) up to the DC term the zero of the covariance removes
    new_delta = M.to_real(M.apply_transfer(new, cov, 0.5))
    d = (delta - new_delta).reshape(-1)[g.flags]
    assert d.std() < 1e-12


def test_interpolation_is_exact_for_cubics_and_deinterpolation_is_its_transpose():
    """Properties of the tricubic pair (tricubic.hpp:54-77,450-549): reproduces polynomials up to degree 2 exactly
    per axis away from the clamped boundary, and <P x, y> == <x, P^T y>."""
    rng = np.random.default_rng(1)
    base = ML.Grid(32.0, 32)
    fine = ML.Grid(32.0, 32, 0.5, (7.0, 9.0, 4.0))
    x = rng.standard_normal((32, 32, 32))
    y = rng.standard_normal((32, 32, 32))
    Px = ML.interpolate_from(fine, base, x)
    PTy = ML.deinterpolate_into(base, fine, y, 1.0)
    assert abs(np.sum(Px * y) - np.sum(x * PTy)) < 1e-10 * abs(np.sum(Px * y))
    c = [base.centroids_1d(a) for a in range(3)]
    q = 0.3 * c[0][:, None, None] ** 2 - c[1][None, :, None] * c[2][None, None, :] + 2.0 * c[2][None, None, :]
    cf = [fine.centroids_1d(a) for a in range(3)]
    qf = 0.3 * cf[0][:, None, None] ** 2 - cf[1][None, :, None] * cf[2][None, None, :] + 2.0 * cf[2][None, None, :]
    assert np.abs(ML.interpolate_from(fine, base, q) - qf).max() < 1e-10

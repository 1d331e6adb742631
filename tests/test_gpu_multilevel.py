"""GPU parity for the zoom-level part of the path: cross-grid transfers, windowed metric,
multi-level covectors, linear / variance modifications and splicing, all through the C ABI, against the CPU oracle
(oracle/multilevel.py, itself pinned on the reference's goldens in test_oracle_multilevel.py) and against the same
goldens directly. Tolerance: 1e-10 relative to the field's scale in double precision (north star)."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import genetic_b200 as gb
from genetic_b200 import multilevel as GM
from oracle import model as M, multilevel as ML

TOL = 1e-10


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _camb_pair(golden_dir):
    fn = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    dev = {s: gb.CAMB(cos, fn, s) for s in ("dm", "all")}
    return cos, dev, M.CAMB(fn, M.Cosmology(0.279, 0.721, 0.817, 99))


def _context(ogrids, spectra=None):
    """Device-side MultiLevelGrid mirroring a list of oracle grids (flags included)."""
    ctx = GM.MultiLevelGrid()
    for og in ogrids:
        g = gb.Grid(og.simsize, og.n, dx=og.cell, x0=og.offset[0], y0=og.offset[1], z0=og.offset[2])
        g.flagCells(og.flags)
        ctx.addLevel(g)
    for k, v in (spectra or {}).items():
        ctx.setPowerspectrumGenerator(k, v)
    return ctx


def _upload(ctx, arrays, fourier=True, **kw):
    return GM.MultiLevelField(ctx, [gb.Field(g, a, fourier=fourier) for g, a in zip(ctx.grids, arrays)], **kw)


@pytest.mark.parametrize("ratio,n_fine,corner", [(2, 32, (7, 9, 4)), (4, 64, (0, 20, 31)), (2, 64, (0, 0, 0))])
def test_interpolate_subsample_deinterpolate_vs_oracle(ratio, n_fine, corner):
    rng = np.random.default_rng(3)
    base = ML.Grid(32.0, 32)
    fine = ML.Grid(32.0, n_fine, 1.0 / ratio, tuple(float(c) for c in corner))
    gbase = gb.Grid(32.0, 32)
    gfine = gb.Grid(32.0, n_fine, dx=1.0 / ratio, x0=corner[0], y0=corner[1], z0=corner[2])
    x = rng.standard_normal((32, 32, 32))
    y = rng.standard_normal((n_fine,) * 3)
    # coarse -> fine (tricubic)
    f = gb.Field(gfine, y, fourier=False)
    f.addFieldFromDifferentGrid(gb.Field(gbase, x, fourier=False))
    assert _rel(f.getDataVector(), y + ML.interpolate_from(fine, base, x)) < TOL
    # fine -> coarse (sub-cell mean)
    c = gb.Field(gbase, x, fourier=False)
    c.addFieldFromDifferentGrid(gb.Field(gfine, y, fourier=False))
    assert _rel(c.getDataVector(), x + ML.subsample_from(base, fine, y)) < TOL
    # conjugate deinterpolation
    d = gb.Field(gbase, x, fourier=False)
    src = gb.Field(gfine, y, fourier=False)
    d.lib.gx_deinterp_add(d.ptr, GM.L.C.byref(gbase.desc()), src.ptr, GM.L.C.byref(gfine.desc()), d.dt, 0.125, None)
    assert _rel(d.getDataVector(), x + ML.deinterpolate_into(base, fine, y, 0.125)) < TOL


def test_partial_coarse_cells_and_clamped_parent():
    """zoom factor 3 with 64 cells covers 21 1/3 coarse cells (virtualgrid.hpp:570-640), and a parent that does not span
    the box clamps its interpolation stencil (field.hpp:598-617)."""
    rng = np.random.default_rng(4)
    base, gbase = ML.Grid(64.0, 64), gb.Grid(64.0, 64)
    mid = ML.Grid(64.0, 64, 1.0 / 3.0, (22.0, 22.0, 22.0))
    gmid = gb.Grid(64.0, 64, dx=1.0 / 3.0, x0=22.0, y0=22.0, z0=22.0)
    x, y = rng.standard_normal((64,) * 3), rng.standard_normal((64,) * 3)
    c = gb.Field(gbase, x, fourier=False)
    c.addFieldFromDifferentGrid(gb.Field(gmid, y, fourier=False))
    assert _rel(c.getDataVector(), x + ML.subsample_from(base, mid, y)) < TOL
    # second zoom touching the edge of the first: clamped stencil
    off = 22.0 + 32 * (1.0 / 3.0)
    fin = ML.Grid(64.0, 64, 1.0 / 6.0, (off,) * 3)
    gfin = gb.Grid(64.0, 64, dx=1.0 / 6.0, x0=off, y0=off, z0=off)
    f = gb.Field(gfin, x, fourier=False)
    f.addFieldFromDifferentGrid(gb.Field(gmid, y, fourier=False))
    assert _rel(f.getDataVector(), x + ML.interpolate_from(fin, mid, y)) < TOL


def test_windows_and_masks_vs_oracle():
    rng = np.random.default_rng(5)
    base, gbase = ML.Grid(64.0, 64), gb.Grid(64.0, 64)
    for off in ((8.0, 15.0, 7.0), (40.0, 0.0, 50.0)):
        win, gwin = ML.Grid(64.0, 64, 0.5, off), gb.Grid(64.0, 64, dx=0.5, x0=off[0], y0=off[1], z0=off[2])
        x = rng.standard_normal((64,) * 3)
        m = ML.window_mask(base, win)
        f = gb.Field(gbase, x, fourier=False)
        f.setZeroInsideWindow(gwin)
        assert np.array_equal(f.getDataVector(), np.where(m, 0.0, x))
        f = gb.Field(gbase, x, fourier=False)
        f.setZeroOutsideWindow(gwin)
        assert np.array_equal(f.getDataVector(), np.where(m, x, 0.0))
    ML.select_sphere([base], (30, 31, 33), 9)
    gbase.flagCells(base.flags)
    f = gb.Field(gbase, x, fourier=False)
    s = GM.L.C.c_double()
    f.lib.gx_masked_sum(f.ptr, gbase.flagMask(), 64, f.dt, GM.L.C.byref(s), None)
    assert abs(s.value - x.reshape(-1)[base.flags].sum()) < 1e-12 * len(base.flags)
    f.lib.gx_masked_shift_scale(f.ptr, gbase.flagMask(), 64, f.dt, 0.25, 3.0, None)
    ref = np.zeros(64 ** 3)
    ref[base.flags] = (x.reshape(-1)[base.flags] - 0.25) * 3.0
    assert np.array_equal(f.getDataVector().reshape(-1), ref)


def test_chi2_two_levels_golden_02m(golden_dir):
    og = [ML.Grid(64.0, 64)]
    ML.select_cube(og, (25, 32, 24), 16)
    ML.add_zoom_level(og, 2, 64)
    ctx = _context(og)
    field = GM.draw(ctx, 13842314)          # random_seed_real_space: one gsl stream across the two levels
    text = open(os.path.join(golden_dir, "test_02m_chisquare_twolevels", "reference.txt")).read()
    ref, dof = re.search(r"chi\^2 = ([0-9.]+) \(dof = (\d+)\)", text).groups()
    assert ctx.getNumDof() == int(dof)
    assert "%.10g" % field.getChi2() == ref


def test_metric_vs_oracle_three_levels():
    og = [ML.Grid(64.0, 64)]
    ML.select_cube(og, (30, 30, 30), 20)
    ML.add_zoom_level(og, 2, 64)
    ML.select_cube(og, (30, 30, 30), 8)
    ML.add_zoom_level(og, 2, 64)
    ctx = _context(og)
    ofields = ML.draw_levels(og, 77, "fourier_serial")
    for to_cov in (True, False):
        f = _upload(ctx, ofields, isCovector=not to_cov)
        f.convertToCovector() if to_cov else f.convertToVector()
        ref = ML.apply_metric(og, ofields, to_cov)
        ind = [M.independent_mask(g.n) for g in og]
        for l in range(3):
            assert _rel(f.getFieldForLevel(l).getDataVector()[ind[l]], ref[l][ind[l]]) < TOL


def test_zoom_combination_vs_oracle_and_golden_01c(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = [ML.Grid(50.0, 64)]
    og[0].flags = np.sort(np.loadtxt(os.path.join(golden_dir, "test_01c_zoom_grid", "zoomtest.txt"), dtype=np.int64))
    ML.add_zoom_level(og, 3, 64)
    ctx = _context(og, dev)
    field = GM.draw(ctx, 8896131)
    field.applyPowerSpectrumFor(GM.ALL)
    gen = GM.MultiLevelParticleGenerator(field)
    ofields = ML.draw_levels(og, 8896131, "real")
    odelta = [M.apply_transfer(f, ocamb.covariance(g.size, g.n, "all"), 0.5) for f, g in zip(ofields, og)]
    ref = ML.combine_levels(og, odelta)
    z = np.load(os.path.join(golden_dir, "test_01c_zoom_grid", "grids_every4.npz"))
    for l in range(2):
        d = field.getFieldForLevel(l)
        d.toReal()
        d = d.getDataVector()
        assert _rel(d, ref[l][0]) < TOL
        assert np.abs(d[::4, ::4, ::4] - z["grid%d" % l]).max() < 1e-6 * z["grid%d" % l].std()
        for a in range(3):
            assert _rel(gen.getGeneratorForLevel(l).getGeneratedFields()[a].getDataVector(), ref[l][1][a]) < TOL


def test_two_linear_modifications_golden_02c(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = [ML.Grid(64.0, 64)]
    og[0].flags = np.sort(np.loadtxt(os.path.join(golden_dir, "test_02c_two_modifs", "zoomtest.txt"), dtype=np.int64).reshape(-1))
    ML.add_zoom_level(og, 3, 64)
    ctx = _context(og, dev)
    field = GM.draw(ctx, 8896131)
    mm = GM.ModificationManager(ctx, cos, field)
    ofield = ML.draw_levels(og, 8896131, "real")
    cdm = [ocamb.covariance(g.size, g.n, "dm") for g in og]
    for centre, rel in (((32.5, 31.5, 32.5), 10.0), ((32.5, 33.5, 32.5), -10.0)):
        ML.select_nearest(og, centre)
        for g, o in zip(ctx.grids, og):
            g.flagCells(o.flags)
        v = mm.calculateCurrentValueByName("overdensity")
        assert abs(v / ML.inner_product(og, ML.overdensity_covector(og, cdm), ofield) - 1) < 1e-9
        mm.addModificationToList("overdensity", "relative", rel)
    mm.applyModifications()
    field.applyPowerSpectrumFor(GM.ALL)
    GM.MultiLevelParticleGenerator(field)
    z = np.load(os.path.join(golden_dir, "test_02c_two_modifs", "grids_every4.npz"))
    for l in range(2):
        d = field.getFieldForLevel(l)
        d.toReal()
        ref = z["grid%d" % l]
        assert np.abs(d.getDataVector()[::4, ::4, ::4] - ref).max() < 1e-8 * ref.std()


def test_velocity_covector_vs_oracle(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = [ML.Grid(64.0, 64)]
    ML.select_cube(og, (20, 20, 20), 16)
    ML.add_zoom_level(og, 2, 64)
    ML.select_sphere(og, (20, 20, 20), 3)
    ctx = _context(og, dev)
    mod = GM.VelocityModification(ctx, cos, 1)
    cov = mod.getCovector(GM.WHITENOISE)
    cdm = [ocamb.covariance(g.size, g.n, "dm") for g in og]
    ref = ML.velocity_covector(og, cdm, 1, M.velocity_to_offset_ratio(M.Cosmology(0.279, 0.721, 0.817, 99)))
    for l in range(2):
        ind = M.independent_mask(og[l].n)
        assert _rel(cov.getFieldForLevel(l).getDataVector()[ind], ref[l][ind]) < 1e-9


def test_velocity_angular_momentum_potential_modifications_vs_oracle(golden_dir):
    """`modify vx absolute 0` (test_14), `modify Ly absolute 100` (test_22) and a potential modification through the
    ModificationManager against the oracle (which is pinned on the gadget goldens of test_14 / test_22)."""
    from genetic_b200 import zoom as Z
    cos, dev, ocamb = _camb_pair(golden_dir)
    ocos = M.Cosmology(0.279, 0.721, 0.817, 99)
    ratio = M.velocity_to_offset_ratio(ocos)
    cases = []
    og = ML.Grid(50.0, 8)
    og.flags = np.sort(np.loadtxt(os.path.join(golden_dir, "test_14_velocity_modif", "ids.txt"), dtype=np.int64))
    cases.append((og, "vx", 0.0, lambda g, c: ML.velocity_covector([g], c, 0, ratio)))
    og = ML.Grid(50.0, 32)
    ML.select_ellipse([og], (25, 25, 25), 10, 7.5, 5)
    cases.append((og, "Ly", 100.0, lambda g, c: ML.angular_momentum_covector([g], c, 1, ratio)))
    og = ML.Grid(50.0, 32)
    ML.select_sphere([og], (20, 30, 25), 6)
    cases.append((og, "potential", 1e-6, lambda g, c: ML.potential_covector([g], c, 0.279, 0.01)))
    for og, name, target, ocov in cases:
        ctx = _context([og], dev)
        field = GM.draw(ctx, 8896131)
        mm = GM.ModificationManager(ctx, cos, field)
        cdm = [ocamb.covariance(og.size, og.n, "dm")]
        of = ML.draw_levels([og], 8896131, "real")
        a = ocov(og, cdm)
        v = mm.calculateCurrentValueByName(name)
        assert abs(v - ML.inner_product([og], a, of)) < 1e-9 * max(abs(v), 1e-30)
        mm.addModificationToList(name, "absolute", target)
        mm.applyModifications()
        of2, _ = ML.apply_linear_modifications([og], of, [a], [target])
        ind = M.independent_mask(og.n)
        assert _rel(field.getFieldForLevel(0).getDataVector()[ind], of2[0][ind]) < 1e-9
        assert abs(mm.calculateCurrentValueByName(name) - target) < 1e-9 * max(abs(target), abs(v))


def test_variance_values_golden_02f_02g(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = [ML.Grid(64.0, 64)]
    ML.select_cube(og, (8, 8, 8), 20)
    ctx = _context(og, dev)
    mm = GM.ModificationManager(ctx, cos, GM.draw(ctx, 8896131))
    ref = open(os.path.join(golden_dir, "test_02f_onelevelvariance", "reference.txt")).read()
    assert "variance: calculated value = %.6g" % mm.calculateCurrentValueByName("variance", filterscale=10000.) in ref

    og = [ML.Grid(64.0, 64)]
    ML.select_cube(og, (8, 8, 8), 16)
    ML.add_zoom_level(og, 4, 64, corner=(0, 0, 0))
    ML.select_cube(og, (8, 8, 8), 16)
    ctx = _context(og, dev)
    mm = GM.ModificationManager(ctx, cos, GM.draw(ctx, 8896131))
    ref = open(os.path.join(golden_dir, "test_02g_multilevelvariance", "reference.txt")).read()
    assert "overdensity: calculated value = %.6g" % mm.calculateCurrentValueByName("overdensity") in ref
    assert "variance: calculated value = %.6g" % mm.calculateCurrentValueByName("variance", filterscale=10000.) in ref


def test_variance_modification_reaches_target_and_matches_oracle(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = [ML.Grid(50.0, 32)]
    ML.select_sphere(og, (25, 25, 25), 8)
    ctx = _context(og, dev)
    field = GM.draw(ctx, 424242)
    mm = GM.ModificationManager(ctx, cos, field)
    v0 = mm.calculateCurrentValueByName("variance", filterscale=20.)
    mm.addModificationToList("variance", "relative", 1.5, filterscale=20., initNumberSteps=4, precision=0.01)
    mm.applyModifications()
    v1 = mm.calculateCurrentValueByName("variance", filterscale=20.)
    assert abs(v1 / (1.5 * v0) - 1) < 0.01
    of = ML.draw_levels(og, 424242, "real")
    cdm = [ocamb.covariance(g.size, g.n, "dm") for g in og]
    assert abs(v0 / ML.variance_value(og, cdm, of, 20.) - 1) < 1e-9


def test_overdensity_and_variance_modification_golden_02i(golden_dir):
    """The reference's test_02i through the host layer: selection, zoom placement, both modifications, level combination."""
    from genetic_b200 import zoom as Z
    cos, dev, ocamb = _camb_pair(golden_dir)
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(64.0, 64))
    for k, v in dev.items():
        ctx.setPowerspectrumGenerator(k, v)
    Z.selectSphere(ctx, (32.5, 32.5, 32.5), 5)
    Z.initZoomGrid(ctx, 4, 64)
    Z.selectCube(ctx, (32.5, 32.5, 32.5), 6.0)
    field = GM.draw(ctx, 8896131)
    mm = GM.ModificationManager(ctx, cos, field)
    mm.addModificationToList("overdensity", "relative", 1.0)
    mm.addModificationToList("variance", "relative", 2.0, filterscale=2.0)
    mm.applyModifications()
    field.applyPowerSpectrumFor(GM.ALL)
    GM.MultiLevelParticleGenerator(field)
    z = np.load(os.path.join(golden_dir, "test_02i_multilevel_algorithm", "grids_every4.npz"))
    for l in range(2):
        d = field.getFieldForLevel(l)
        d.toReal()
        ref = z["grid%d" % l]
        assert np.abs(d.getDataVector()[::4, ::4, ::4] - ref).max() < 1e-6 * ref.std()


def test_example_paramfile_fields_vs_oracle(golden_dir):
    """BASELINE.json config 1 (example/paramfile.txt as shipped): Planck-2018 transfer, base 128^3 over 50 Mpc/h, zoom x3
    128^3 around (32.5, 32.5, 32.5), random_seed 8896131 (parallel Fourier). Overdensity and displacement fields of both
    levels after the level combination against the CPU oracle (the particle mapper and the gadget writer are the
    `next` rows of SURVEY 8f)."""
    from genetic_b200 import zoom as Z
    fn = os.path.join(golden_dir, "planck_2018_transfer_out.dat")
    kw = dict(ns=0.9660, hubble=0.6732)
    cos = gb.CosmologicalParameters(0.3158, 0.6842, 0.8117, 200, **kw)
    ocamb = M.CAMB(fn, M.Cosmology(0.3158, 0.6842, 0.8117, 200, **kw))
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(50.0, 128))
    ctx.setPowerspectrumGenerator("all", gb.CAMB(cos, fn, "all"))
    Z.selectSphere(ctx, (32.5, 32.5, 32.5), 6)
    Z.initZoomGrid(ctx, 3, 128)
    og = [ML.Grid(50.0, 128)]
    ML.select_sphere(og, (32.5, 32.5, 32.5), 6)
    ML.add_zoom_level(og, 3, 128)
    assert tuple(float(v) for v in ctx.grids[1].offsetLower) == og[1].offset
    field = GM.draw(ctx, 8896131, drawInFourierSpace=True, parallel=True)
    field.applyPowerSpectrumFor(GM.ALL)
    gen = GM.MultiLevelParticleGenerator(field)
    ofields = ML.draw_levels(og, 8896131, "fourier_parallel")
    ref = ML.combine_levels(og, [M.apply_transfer(f, ocamb.covariance(g.size, g.n, "all"), 0.5) for f, g in zip(ofields, og)])
    for l in range(2):
        d = field.getFieldForLevel(l)
        d.toReal()
        assert _rel(d.getDataVector(), ref[l][0]) < TOL
        for a in range(3):
            assert _rel(gen.getGeneratorForLevel(l).getGeneratedFields()[a].getDataVector(), ref[l][1][a]) < TOL


def test_splice_golden_24_and_oracle(golden_dir):
    cos, dev, ocamb = _camb_pair(golden_dir)
    og = ML.Grid(50.0, 16)
    ML.select_sphere([og], (25, 25, 25), 10)
    ctx = _context([og], dev)
    g = ctx.getGridForLevel(0)
    orig = GM.draw(ctx, 8896131).getFieldForLevel(0)
    new = GM.draw(ctx, 8896132).getFieldForLevel(0)
    cov = dev["all"].getPowerSpectrumForGrid(g)
    sp = GM.spliceOneLevel(new, orig, cov)
    sp.toFourier()
    sp.applyTransferFunction(dev["all"], 0.5)
    sp.toReal()
    delta = sp.getDataVector()
    ref = np.load(os.path.join(golden_dir, "test_24_splicing", "grid0.npz"))["grid0"]
    assert np.abs(delta - ref).max() < 1e-5 * ref.std()
    ocov = ocamb.covariance(og.size, og.n, "all")
    osp = ML.splice_one_level(og, ML.draw_levels([og], 8896132, "real")[0], ML.draw_levels([og], 8896131, "real")[0], ocov)
    odelta = M.to_real(M.apply_transfer(osp, ocov, 0.5))
    assert np.abs(delta - odelta).max() < 1e-5 * ref.std()   # both stop CG at rtol 1e-6

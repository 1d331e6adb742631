"""Particle mappers, host side (pure index arithmetic, no device): id files through a nested TwoLevelParticleMapper
against the reference's mapper_test_01a golden and against the reference binary itself on a three-level zoom."""
import os

import numpy as np
import pytest

import oracle
import genetic_b200 as gb
from genetic_b200 import lib as L, multilevel as GM, zoom as Z, mapper as PM


def _ids(fn):
    return np.loadtxt(fn, dtype=np.int64).reshape(-1)


def test_idfile_round_trip_golden_mapper_test_01a(golden_dir):
    """mapper_test_01a_zoom: ids of a 64^3 base grid define the zoom; reference.txt lists particle ids of the zoomed
    mapper just before / just after the first replaced particle, after all of them and inside the high-res block; the
    reference's test is `idfile reference.txt; dump_idfile output.txt; diff reference.txt output.txt`."""
    d = os.path.join(golden_dir, "mapper_test_01a_zoom")
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(50.0, 64))
    m0 = PM.updateParticleMapper(ctx)
    m0.flagParticles(np.unique(_ids(os.path.join(d, "zoomtest.txt"))))
    Z.initZoomGrid(ctx, 3, 64)
    mapper = PM.updateParticleMapper(ctx)
    assert mapper.size() == 64 ** 3 + (27 - 1) * 1081
    ref = _ids(os.path.join(d, "reference.txt"))
    mapper.unflagAllParticles()
    mapper.flagParticles(np.unique(ref))
    assert np.array_equal(mapper.getFlaggedParticles(), ref)
    # the special values the reference's paramfile comments on
    lev, cell = mapper.layout()
    base, zoom = mapper.pLevel1.layer, mapper.pLevel2.layer
    assert (lev[108381], cell[108381]) == (base, 108381) and (lev[108382], cell[108382]) == (base, 108386)
    assert cell[200000] == 200000 + 1081 and lev[276828] == zoom


def test_mapper_errors():
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(50.0, 16))
    with pytest.raises(L.GxError, match="without marking particles"):
        Z.initZoomGrid(ctx, 2, 16)
    Z.selectSphere(ctx, (25, 25, 25), 6)
    Z.initZoomGrid(ctx, 2, 16)
    m = PM.updateParticleMapper(ctx)
    with pytest.raises(L.GxError, match="ascending order"):
        m.flagParticles([5, 3])
    with pytest.raises(L.GxError, match="out of range"):
        m.flagParticles([m.size()])
    with pytest.raises(L.GxError, match="Consistency error"):      # a cell the level-1 mapper cannot reach
        PM.TwoLevelParticleMapper(m, PM.OneLevelParticleMapper(ctx.getGridForLevel(1), 2), [16 ** 3 + 5])
    with pytest.raises(L.GxError, match="outside high resolution grid"):
        far = np.setdiff1d(np.arange(16 ** 3), ctx.zoomParticleArray[0])[:1]
        PM.TwoLevelParticleMapper(PM.OneLevelParticleMapper(ctx.getGridForLevel(0), 0),
                                  PM.OneLevelParticleMapper(ctx.getGridForLevel(1), 1), far)


@pytest.mark.skipif(oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")
def test_idfile_round_trip_three_levels_vs_reference_binary(tmp_path, golden_dir):
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(60.0, 32))
    Z.selectSphere(ctx, (20, 33, 41), 7)
    Z.initZoomGrid(ctx, 2, 32)
    Z.selectSphere(ctx, (20, 33, 41), 3)
    Z.initZoomGrid(ctx, 2, 32)
    mapper = PM.updateParticleMapper(ctx)
    first1 = mapper.pLevel1.firstLevel2Particle
    rng = np.random.default_rng(5)
    ids = np.unique(np.concatenate([rng.integers(0, mapper.size(), 400), np.arange(first1 - 5, first1 + 40),
                                    np.arange(mapper.size() - 30, mapper.size())]))
    np.savetxt(str(tmp_path / "ids.txt"), ids, fmt="%d")
    pf = ("Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\nrandom_seed_real_space 5\noutname t\n"
          "outdir ./\noutformat tipsy\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\n"
          "centre 20 33 41\nselect_sphere 3\nzoomgrid 2 32\nidfile ids.txt\ndump_idfile out_ids.txt\n")
    oracle.run_ref(pf, str(tmp_path), [os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")])
    mapper.unflagAllParticles()
    mapper.flagParticles(ids)
    assert np.array_equal(mapper.getFlaggedParticles(), _ids(str(tmp_path / "out_ids.txt")))


@pytest.mark.parametrize("test", ["mapper_test_01a_zoom", "mapper_test_01b_zoom_gas", "mapper_test_03a_zoom_linkage",
                                  "mapper_test_03b_zoom_to_coarse", "mapper_test_05_gadget_ptype",
                                  "mapper_test_07_ids_from_two_inputs", "mapper_test_04a_stray",
                                  "mapper_test_04b_stray_linkage", "mapper_test_04c_stray_wrapping",
                                  "mapper_test_06_intermediate_levels"])
def test_reference_mapper_tests_through_the_command_layer(golden_dir, tmp_path, monkeypatch, test):
    """The reference's mapper_test_* directories (run_mapper_tests.sh: run the paramfile, `diff reference.txt output.txt`)
    through genetic_b200.icgen: zoom mappers with and without gas, per-type gadget ordering, sub-sampled inputs, and
    `mapper_relative_to` (ids of another setup, rebuilt by a field-less run of its paramfile), stray particles, adapt_mask
    and autopad. Pure host logic: the
    commands that need the device (`camb`, `done`) are not reached."""
    import shutil
    from genetic_b200 import icgen
    for d in os.listdir(golden_dir):
        if d.startswith("mapper_test_"):
            shutil.copytree(os.path.join(golden_dir, d), tmp_path / d)
    monkeypatch.setattr(icgen.ICGenerator, "setCambDat", lambda self, path: None)     # also in the field-less sub-runs
    gen = icgen.ICGenerator(log=lambda s: None)
    gen.commands["camb"] = (lambda path: None, str)
    gen.commands["done"] = (lambda: None,)
    cwd = os.getcwd()
    os.chdir(tmp_path / test)
    try:
        for line in open("paramfile.txt"):
            if line.strip():
                gen.run_line(line.rstrip("\n"))
        assert open("reference.txt").read().split() == open("output.txt").read().split()
    finally:
        os.chdir(cwd)

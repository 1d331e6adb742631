"""The C++ host layer (include/genetic_b200.hpp) and its paramfile driver (host/genetic_b200_ic.cpp).

CPU: the driver builds with plain g++ against the C ABI and fails loudly (exit 1 + message, the reference's
parser.hpp:202-207 behaviour) when no device is present. GPU: the reference's golden paramfiles test_13 (parallel Fourier
seed) and test_05a (reverse serial Fourier seed) run through the C++ classes reproduce the reference's tipsy files."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from genetic_b200 import build as gbuild
from oracle import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_driver_builds_and_fails_loudly_without_a_device(tmp_path, golden_dir):
    exe = gbuild.build_host()
    assert os.access(exe, os.X_OK)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    if _have_gpu():
        pytest.skip("a device is present")
    r = subprocess.run([exe, os.path.join(golden_dir, "test_13_parallel_seed", "paramfile.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Error:")


@pytest.mark.gpu
@pytest.mark.parametrize("test,name", [("test_13_parallel_seed", "test_13"), ("test_05a_seed_old", "test_5")])
def test_reference_paramfiles_through_the_cpp_host(tmp_path, golden_dir, test, name):
    exe = gbuild.build_host()
    d = tmp_path / test
    d.mkdir()
    shutil.copy(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), tmp_path / "camb_transfer_kmax40_z0.dat")
    shutil.copy(os.path.join(golden_dir, test, "paramfile.txt"), d / "paramfile.txt")
    r = subprocess.run([exe, str(d / "paramfile.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = M.read_tipsy(str(d / (name + ".tipsy")))
    g = M.read_tipsy(os.path.join(golden_dir, test, "reference_output"))
    assert t["n"] == g["n"] and t["a"] == g["a"]
    ulp_p = np.abs(t["pos"] - g["pos"]) / np.spacing(np.abs(g["pos"]))
    ulp_v = np.abs(t["vel"] - g["vel"]) / np.spacing(np.abs(g["vel"]))
    assert ulp_p.max() <= 1 and ulp_v.max() <= 1
    assert (ulp_p == 0).mean() > 0.99 and (ulp_v == 0).mean() > 0.99
    assert np.all(g["mass"] == t["mass"]) and np.all(g["eps"] == t["eps"])


@pytest.mark.gpu
def test_unknown_command_is_an_error(tmp_path):
    exe = gbuild.build_host()
    p = tmp_path / "paramfile.txt"
    p.write_text("Om 0.3\nzoomgrid 2 64\n")
    r = subprocess.run([exe, str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown command zoomgrid" in r.stderr


@pytest.mark.gpu
def test_chi2_goldens_through_the_cpp_multilevel_classes(tmp_path, golden_dir):
    """test_02k (one level) and test_02m (zoom x2) chi^2 through gx::fields::MultiLevelField (C++): the windowed covector
    metric, tricubic / sub-cell transfers and the weighted mode sums; the zoom window comes from genetic_b200/zoom.py."""
    import re
    import genetic_b200 as gb
    from genetic_b200 import multilevel as GM, zoom as Z
    exe = gbuild.build_host(program="genetic_b200_chi2")
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(64.0, 64))
    for test, zoom in (("test_02k_chisquare_onelevel", False), ("test_02m_chisquare_twolevels", True)):
        if zoom:
            Z.selectCube(ctx, (25, 32, 24), 16)
            Z.initZoomGrid(ctx, 2, 64)
        lv = tmp_path / (test + ".txt")
        lv.write_text("".join("%r %d %r %r %r %r\n" % (float(g.periodicDomainSize), g.size, float(g.cellSize),
                                                      *(float(v) for v in g.offsetLower)) for g in ctx.grids))
        r = subprocess.run([exe, str(lv), "13842314"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        ref = open(os.path.join(golden_dir, test, "reference.txt")).read()
        want = re.search(r"Calculated chi\^2 = [0-9.]+ \(dof = \d+\)", ref).group(0)
        assert want in r.stdout, (want, r.stdout)

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

"""The reference ITSELF (oracle/_ref/genetIC: the unmodified sources of /root/reference compiled against the stand-in
FFTW3 / GSL / HDF5 headers of oracle/shims/, see oracle/build_ref.sh) run on the reference's golden paramfiles and on NEW
inputs, against (a) its own goldens and (b) the numpy oracle that the GPU parity tests use. CPU only; needs the prebuilt
binary (built by __graft_entry__.build() where /root/reference exists) but not /root/reference at run time."""
import os
import re

import numpy as np
import pytest

import oracle
from oracle import model as M, multilevel as ML

pytestmark = pytest.mark.skipif(oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built (no /root/reference here)")
COSMO = "Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\n"


def _cosmo():
    return M.Cosmology(0.279, 0.721, 0.817, 99)


def _camb_file(golden_dir):
    return os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")


def test_reference_binary_reproduces_its_goldens(tmp_path, golden_dir):
    """test_13 tipsy bit for bit, test_02m chi^2 text: the shims are transparent."""
    pf = open(os.path.join(golden_dir, "test_13_parallel_seed", "paramfile.txt")).read().replace("../camb", "camb")
    oracle.run_ref(pf, str(tmp_path / "t13"), [_camb_file(golden_dir)])
    a = M.read_tipsy(str(tmp_path / "t13" / "test_13.tipsy"))
    b = M.read_tipsy(os.path.join(golden_dir, "test_13_parallel_seed", "reference_output"))
    for k in ("mass", "pos", "vel", "eps"):
        assert np.array_equal(a[k], b[k])
    pf = open(os.path.join(golden_dir, "test_02m_chisquare_twolevels", "paramfile.txt")).read().replace("../camb", "camb")
    out = oracle.run_ref(pf, str(tmp_path / "t02m"), [_camb_file(golden_dir)])
    ref = open(os.path.join(golden_dir, "test_02m_chisquare_twolevels", "reference.txt")).read().strip()
    assert ref in out


@pytest.mark.parametrize("seedcmd,mode,n,box,seed", [
    ("random_seed", "fourier_parallel", 32, 75.0, 20261017), ("random_seed_serial", "fourier_serial", 16, 40.0, 777),
    ("random_seed_real_space", "real", 32, 120.0, 31337)])
def test_oracle_unigrid_matches_the_reference_on_new_inputs(tmp_path, golden_dir, seedcmd, mode, n, box, seed):
    """Seeds, sizes and boxes that appear in none of the goldens: overdensity grid (float64 npy) to 1e-12 sigma and the
    tipsy particles to float32 rounding."""
    pf = COSMO + "%s %d\noutname t\noutdir ./\noutformat tipsy\nbasegrid %r %d\ndone\ndump_grid 0\n" % (seedcmd, seed, box, n)
    oracle.run_ref(pf, str(tmp_path), [_camb_file(golden_dir)])
    r = M.unigrid_ic(n, box, seed, mode, _camb_file(golden_dir), _cosmo(), return_intermediate=True)
    grid = np.load(str(tmp_path / "grid-0.npy"))
    assert np.abs(M.to_real(r["delta_k"]) - grid).max() < 1e-12 * grid.std()
    t = M.to_tipsy_units(r["pos"], r["vel"], r["mass"], r["eps"], box, _cosmo(), n ** 3)
    g = M.read_tipsy(str(tmp_path / "t.tipsy"))
    assert np.abs(t["pos"] - g["pos"]).max() <= np.spacing(np.float32(0.5)) and (t["pos"] == g["pos"]).mean() > 0.99
    assert np.abs(t["vel"] - g["vel"]).max() <= 2 * np.spacing(np.abs(g["vel"]).max()) and (t["vel"] == g["vel"]).mean() > 0.99
    assert np.all(g["mass"] == t["mass"]) and np.all(g["eps"] == t["eps"])


def test_oracle_zoom_modifications_match_the_reference_on_new_inputs(tmp_path, golden_dir):
    """A new two-level configuration (zoom x2 around an off-centre sphere, another seed) with an overdensity and a velocity
    modification: chi^2, calculated values and the per-level grids after `done`."""
    pf = COSMO + ("random_seed_real_space 424243\noutname t\noutdir ./\noutformat tipsy\nbasegrid 60.0 32\n"
                  "centre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\ncentre 20 33 41\nselect_sphere 4\n"
                  "chi2\ncalculate overdensity\ncalculate vy\nmodify overdensity absolute 0.05\nmodify vy relative 0.5\n"
                  "done\ndump_grid 0\ndump_grid 1\n")
    out = oracle.run_ref(pf, str(tmp_path), [_camb_file(golden_dir)])
    camb = M.CAMB(_camb_file(golden_dir), _cosmo())
    grids = [ML.Grid(60.0, 32)]
    ML.select_sphere(grids, (20, 33, 41), 7)
    ML.add_zoom_level(grids, 2, 32)
    ML.select_sphere(grids, (20, 33, 41), 4)
    f = ML.draw_levels(grids, 424243, "real")
    cdm = [camb.covariance(g.size, g.n, "dm") for g in grids]
    call = [camb.covariance(g.size, g.n, "all") for g in grids]
    chi2 = float(re.search(r"chi\^2 = ([0-9.]+)", out).group(1))
    assert abs(ML.chi2(grids, f) / chi2 - 1) < 1e-9
    a1 = ML.overdensity_covector(grids, cdm)
    a2 = ML.velocity_covector(grids, cdm, 1, M.velocity_to_offset_ratio(_cosmo()))
    v1, v2 = ML.inner_product(grids, a1, f), ML.inner_product(grids, a2, f)
    assert abs(v1 / float(re.search(r"overdensity: calculated value = (\S+)", out).group(1)) - 1) < 1e-5
    assert abs(v2 / float(re.search(r"vy: calculated value = (\S+)", out).group(1)) - 1) < 1e-5
    # ICGenerator::modify(string, string, float target) (ic.hpp:1613): the paramfile's target passes through a float
    f2, _ = ML.apply_linear_modifications(grids, f, [a1, a2], [float(np.float32(0.05)), 0.5 * v2])
    res = ML.combine_levels(grids, [M.apply_transfer(x, c, 0.5) for x, c in zip(f2, call)])
    for l in range(2):
        ref = np.load(str(tmp_path / ("grid-%d.npy" % l)))
        assert np.abs(res[l][0] - ref).max() < 1e-10 * ref.std()

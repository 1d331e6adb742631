"""GPU path against the reference ITSELF: oracle/_ref/genetIC (the unmodified reference sources compiled against stand-in
FFTW3/GSL/HDF5 headers, oracle/build_ref.sh) is run on the GPU box's host cores on inputs that appear in none of the
goldens, and the device results are compared with its output files. f64: 1e-10 relative (north star); particles in the
reference's float32 tipsy file to float32 rounding."""
import os
import re

import numpy as np
import pytest

import oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built (no /root/reference at build time)")]

import genetic_b200 as gb
from genetic_b200 import multilevel as GM, zoom as Z
from oracle import model as M

COSMO = "Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\n"


def _camb_file(golden_dir):
    return os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")


@pytest.mark.parametrize("seedcmd,mode,n,box,seed", [
    ("random_seed", "fourier_parallel", 64, 75.0, 20261017), ("random_seed_serial", "fourier_serial", 32, 40.0, 777),
    ("random_seed_real_space", "real", 64, 120.0, 31337)])
def test_unigrid_ic_matches_the_reference_binary(tmp_path, golden_dir, seedcmd, mode, n, box, seed):
    pf = COSMO + "%s %d\noutname t\noutdir ./\noutformat tipsy\nbasegrid %r %d\ndone\ndump_grid 0\n" % (seedcmd, seed, box, n)
    oracle.run_ref(pf, str(tmp_path), [_camb_file(golden_dir)])
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    # overdensity field through the class surface
    grid = gb.Grid(box, n)
    field = gb.Field(grid, fourier=True)
    rfg = gb.RandomFieldGenerator(field)
    rfg.seed(seed)
    rfg.setDrawInFourierSpace(mode != "real")
    rfg.setParallel(mode == "fourier_parallel")
    rfg.draw()
    field.applyTransferFunction(gb.CAMB(cos, _camb_file(golden_dir), "all"), 0.5)
    field.toReal()
    ref = np.load(str(tmp_path / "grid-0.npy"))
    assert np.abs(field.getDataVector() - ref).max() < 1e-10 * ref.std()
    # particles through the whole chain
    pos, vel, mass, eps = gb.generate_unigrid_ic(n, box, seed, cos, _camb_file(golden_dir), mode=mode, out_dtype=np.float64)
    t = M.to_tipsy_units(pos, vel, mass, eps, box, M.Cosmology(0.279, 0.721, 0.817, 99), n ** 3)
    g = M.read_tipsy(str(tmp_path / "t.tipsy"))
    assert np.abs(t["pos"] - g["pos"]).max() <= np.spacing(np.float32(0.5)) and (t["pos"] == g["pos"]).mean() > 0.99
    assert np.abs(t["vel"] - g["vel"]).max() <= 2 * np.spacing(np.abs(g["vel"]).max()) and (t["vel"] == g["vel"]).mean() > 0.99
    assert np.all(g["mass"] == t["mass"]) and np.all(g["eps"] == t["eps"])


def test_zoom_with_modifications_matches_the_reference_binary(tmp_path, golden_dir):
    pf = COSMO + ("random_seed_real_space 424243\noutname t\noutdir ./\noutformat tipsy\nbasegrid 60.0 32\n"
                  "centre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\ncentre 20 33 41\nselect_sphere 4\n"
                  "chi2\ncalculate overdensity\ncalculate vy\nmodify overdensity absolute 0.05\nmodify vy relative 0.5\n"
                  "done\ndump_grid 0\ndump_grid 1\n")
    out = oracle.run_ref(pf, str(tmp_path), [_camb_file(golden_dir)])
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(60.0, 32))
    for s in ("dm", "all"):
        ctx.setPowerspectrumGenerator(s, gb.CAMB(cos, _camb_file(golden_dir), s))
    Z.selectSphere(ctx, (20, 33, 41), 7)
    Z.initZoomGrid(ctx, 2, 32)
    Z.selectSphere(ctx, (20, 33, 41), 4)
    field = GM.draw(ctx, 424243)
    chi2 = float(re.search(r"chi\^2 = ([0-9.]+)", out).group(1))
    assert abs(field.getChi2() / chi2 - 1) < 1e-9
    mm = GM.ModificationManager(ctx, cos, field)
    assert abs(mm.calculateCurrentValueByName("overdensity") /
               float(re.search(r"overdensity: calculated value = (\S+)", out).group(1)) - 1) < 1e-5
    assert abs(mm.calculateCurrentValueByName("vy") / float(re.search(r"vy: calculated value = (\S+)", out).group(1)) - 1) < 1e-5
    mm.addModificationToList("overdensity", "absolute", float(np.float32(0.05)))  # ICGenerator::modify takes a float target
    mm.addModificationToList("vy", "relative", 0.5)
    mm.applyModifications()
    field.applyPowerSpectrumFor(GM.ALL)
    GM.MultiLevelParticleGenerator(field)
    for l in range(2):
        d = field.getFieldForLevel(l)
        d.toReal()
        ref = np.load(str(tmp_path / ("grid-%d.npy" % l)))
        assert np.abs(d.getDataVector() - ref).max() < 1e-9 * ref.std()

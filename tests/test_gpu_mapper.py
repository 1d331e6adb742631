"""Particle mappers + gadget writer against the reference's OWN goldens (test_10b: zoom x3 with variable masses;
test_10c: the zoom level as gadget type 0, the base grid as type 1): every 53rd particle of the reference's
reference_output files (tests/golden/make_golden.py), f64 fields, float32 file. The zoom region comes from the
reference's id file through the mapper's flagParticles, the zoom window from zoom.initZoomGrid."""
import os

import numpy as np
import pytest

import genetic_b200 as gb
from genetic_b200 import multilevel as GM, zoom as Z, mapper as PM, output as OUT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("test,types", [("test_10b_gadget_zoom_varimass", [1, 1]), ("test_10c_gadget_zoom_ptype", [1, 0])])
def test_gadget_zoom_goldens_10b_10c(tmp_path, golden_dir, test, types):
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(50.0, 64))
    for s in ("dm", "all"):
        ctx.setPowerspectrumGenerator(s, gb.CAMB(cos, camb, s))
    ids = np.unique(np.loadtxt(os.path.join(golden_dir, test, "zoomtest.txt"), dtype=np.int64))
    m0 = PM.updateParticleMapper(ctx)
    m0.unflagAllParticles()
    m0.flagParticles(ids)                                   # idfile zoomtest.txt
    assert np.array_equal(m0.getFlaggedParticles(), ids)    # dump_idfile zoomtest_crosscheck.txt
    Z.initZoomGrid(ctx, 3, 64)                              # zoomgrid 3 64
    field = GM.draw(ctx, 8896131)                           # random_seed_real_space 8896131
    field.applyPowerSpectrumFor(GM.ALL)
    gen = GM.MultiLevelParticleGenerator(field)
    mapper = PM.updateParticleMapper(ctx, types)
    parts = PM.MappedParticles(mapper, gen, cos, out_dtype=np.float32)
    fn = OUT.saveGadget(str(tmp_path / "test_1.gadget"), 50.0, parts, cos, 3)[0]
    ours = OUT.read_gadget(fn)
    gold = np.load(os.path.join(golden_dir, test, "particles_every53.npz"))
    assert ours["header"][:208] == gold["header"].tobytes()   # counts, mass table, cosmology, flags (not the 48 fill bytes)
    idx = gold["index"]
    assert np.array_equal(ours["ids"][idx], gold["ids"])      # bit-exact particle ordering
    assert np.array_equal(ours["pos"][idx], gold["pos"])
    # float32 rounding of an f64 result: one ulp, or (for components near zero) 1e-10 of the field's scale
    dv = np.abs(ours["vel"][idx] - gold["vel"])
    assert np.all((dv <= np.spacing(np.abs(gold["vel"]))) | (dv <= 1e-10 * np.abs(gold["vel"]).max()))
    assert (dv == 0).mean() > 0.99
    if "mass" in gold.files:
        assert np.array_equal(ours["mass"][idx], gold["mass"])
    else:
        assert "mass" not in ours


def test_chunked_evaluation_equals_one_shot(golden_dir):
    """MappedParticles stages bounded chunks on the device (id ranges through gx_emit_component, scattered cells through
    gx_emit_cells): any chunk size gives the same particles, also for an interleaved (non-contiguous) layout."""
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99, OmegaBaryons0=0.04)
    ctx = GM.MultiLevelGrid()
    ctx.addLevel(gb.Grid(40.0, 16))
    ctx.setPowerspectrumGenerator("all", gb.CAMB(cos, camb, "all"))
    Z.selectSphere(ctx, (20, 21, 19), 6)
    Z.initZoomGrid(ctx, 2, 16)
    field = GM.draw(ctx, 77)
    field.applyPowerSpectrumFor(GM.ALL)
    gen = GM.MultiLevelParticleGenerator(field)
    mapper = PM.updateParticleMapper(ctx, None, cos, baryonsOnAllLevels=True, gasFirst=True)
    a = PM.MappedParticles(mapper, gen, cos, out_dtype=np.float64)
    b = PM.MappedParticles(mapper, gen, cos, out_dtype=np.float64, chunk=777)
    for k in ("pos", "vel", "mass", "eps", "gadgetType", "isGas"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    # gas first on all levels: level 0 appears twice in the layout (gas, then dark matter) -> not one contiguous run
    lev = a.level
    assert np.count_nonzero(np.diff(lev) != 0) >= 3
    n0 = 16 ** 3 - len(ctx.zoomParticleArray[0])
    assert np.array_equal(a.pos[:n0], a.pos[mapper.size_gas():mapper.size_gas() + n0])   # same cells, two species
    assert np.all(a.mass[:n0] < a.mass[mapper.size_gas():mapper.size_gas() + n0])

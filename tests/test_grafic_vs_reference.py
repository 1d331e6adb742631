"""Grafic output and refinement masks against the reference binary itself (oracle/_ref/genetIC: the unmodified reference
compiled against stand-in FFTW / GSL headers) on cases the reference's own suite does not hold: a centred zoom that wraps
around the box with a passive-scalar value, sub- and supersampled stacks, three levels centred on a third point, gas, flags
changed after the zoom was opened. Host logic on the numpy mock device (tests/host/mock_device.py), so this runs on CPU; levels,
headers, ids, ic_refmap, ic_pvar and the dumped masks must be identical, the fields agree to 1e-5 (mock FFT against the
reference's)."""
import glob
import importlib.util
import os
import shutil

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("gx_mock_device", os.path.join(HERE, "host", "mock_device.py"))
_mock = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mock)

pytestmark = pytest.mark.skipif(oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")

HEAD = ("Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\nrandom_seed_real_space 8896131\noutdir ./\n"
        "outname t\n")
CASES = {
    "centre_wrap_zoom2": "outformat grafic\nbasegrid 50.0 16\ncentre 2 47 25\nselect_sphere 4\nzoomgrid 2 16\ncenter_output\npvar 0.5\ndone\ndump_mask\n",
    "super2_sub4": "outformat 5\nsubsample 4\nsupersample 2\nbasegrid 50.0 16\ncentre 25 25 25\nselect_sphere 3\nzoomgrid 4 16\ndone\ndump_mask\n",
    "super4_offcentre": "outformat 5\nsupersample 4\nbasegrid 50.0 16\ncentre 30 21 14\nselect_sphere 5\nzoomgrid 2 16\ncenter_grafic_output\npvar 0.25\ndone\ndump_mask\n",
    "flags_after_zoom": "outformat grafic\nbasegrid 50.0 16\ncentre 12 12 40\nselect_cube 9\nzoomgrid 2 16\ncentre 12 12 40\nselect_nearest\ncenter_output\ndone\ndump_mask\n",
    "three_levels": "outformat grafic\nbasegrid 50.0 16\ncentre 25 25 25\nselect_sphere 6\nzoomgrid 2 16\ncentre 25 25 25\nselect_sphere 2\nzoomgrid 4 16\ncentre 24 26 25\ncenter_output\ndone\ndump_mask\n",
    "gas_zoom2": "Ob 0.04\noutformat grafic\nbasegrid 50.0 16\ncentre 30 20 10\nselect_cube 6\nzoomgrid 2 16\ndone\n",
    "unigrid_sub2": "outformat grafic\nsubsample 2\nbasegrid 50.0 16\ncentre 10 10 10\ncenter_output\ndone\ndump_mask\n",
}


@pytest.fixture(autouse=True)
def mock_device(monkeypatch):
    import genetic_b200.icgen as I
    m = _mock.install(monkeypatch)
    monkeypatch.setattr(I, "CAMB", _mock.MockSpectrum)
    return m


def _levels(d):
    return sorted(glob.glob(str(d) + "/*.grafic_*"), key=lambda p: int(p.rsplit("_", 1)[1]))


@pytest.mark.parametrize("case", sorted(CASES))
def test_grafic_files_and_masks_match_the_reference_binary(case, golden_dir, tmp_path):
    from genetic_b200 import icgen, grafic as GF
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    text = HEAD + CASES[case]
    oracle.run_ref(text, str(tmp_path / "ref"), [camb])
    ours = tmp_path / "ours"
    ours.mkdir()
    shutil.copy(camb, ours)
    (ours / "paramfile.txt").write_text(text)
    icgen.run_paramfile(str(ours / "paramfile.txt"), log=lambda s: None)
    ref_levels, our_levels = _levels(tmp_path / "ref"), _levels(ours)
    assert [p.rsplit("_", 1)[1] for p in ref_levels] == [p.rsplit("_", 1)[1] for p in our_levels] and ref_levels
    for a, b in zip(ref_levels, our_levels):
        x, y = GF.read_grafic_level(a), GF.read_grafic_level(b)
        np.testing.assert_allclose(y["header"], x["header"], rtol=2e-7, atol=1e-7)
        for k in x:
            if k == "header":
                continue
            if k in ("ic_particle_ids", "ic_refmap", "ic_pvar_00001"):
                assert np.array_equal(x[k], y[k]), (os.path.basename(a), k)
            else:
                assert np.abs(x[k].astype(np.float64) - y[k]).max() < 1e-5 * max(1.0, np.abs(x[k]).max()), (os.path.basename(a), k)
    masks = sorted(glob.glob(str(tmp_path / "ref" / "mask-?.npy")))
    assert bool(masks) == ("dump_mask" in text)
    for f in masks:
        assert np.array_equal(np.load(f), np.load(str(ours / os.path.basename(f)))), os.path.basename(f)

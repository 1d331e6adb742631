"""A fixed set of randomly drawn paramfiles (tools/fuzz_vs_reference.py: zoom geometry, gas, gadget types, flagged type,
sub/supersampling, strays, autopad, adapt_mask, multi-file, velocity offset, all particle formats) through the Python command
layer on the numpy mock device and through the reference binary itself; the particle files must agree block by block. The
second set forces stray particles below a SECOND zoom level (ResolutionMatchingGrid output grids nested twice, ic.hpp:398-479
relative to the output grid above), which none of the reference's own tests exercises."""
import importlib.util
import os
import random

import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("gx_fuzz", os.path.join(os.path.dirname(HERE), "tools", "fuzz_vs_reference.py"))
fuzz = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(fuzz)
_mspec = importlib.util.spec_from_file_location("gx_mock_device", os.path.join(HERE, "host", "mock_device.py"))
_mock = importlib.util.module_from_spec(_mspec)
_mspec.loader.exec_module(_mock)

pytestmark = pytest.mark.skipif(oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")


@pytest.fixture(autouse=True)
def mock_device(monkeypatch):
    import genetic_b200.icgen as I
    m = _mock.install(monkeypatch)
    monkeypatch.setattr(I, "CAMB", _mock.MockSpectrum)
    return m


@pytest.mark.parametrize("seed,strays2,cases,at_least", [(2, False, 24, 10), (7, True, 40, 8)])
def test_random_paramfiles_match_the_reference_binary(tmp_path, seed, strays2, cases, at_least):
    rng = random.Random(seed)
    compared = 0
    for it in range(cases):
        body = fuzz.make_case(rng, strays2)
        diffs = fuzz.compare(body, str(tmp_path / ("case%d" % it)))
        if diffs is None:            # the reference itself refuses this geometry
            continue
        compared += 1
        assert not diffs, (diffs[:4], body)
    assert compared >= at_least


def test_random_grafic_paramfiles_match_the_reference_binary(tmp_path):
    """Grafic output and dump_mask on random stacks (sub/supersampling, 0-2 zooms, pvar, centred output on a moved point)."""
    rng = random.Random(3)
    compared = 0
    for it in range(24):
        body = fuzz.make_grafic_case(rng)
        diffs = fuzz.compare_grafic(body, str(tmp_path / ("g%d" % it)))
        if diffs is None:
            continue
        compared += 1
        assert not diffs, (diffs[:4], body)
    assert compared >= 10


def test_random_modification_paramfiles_match_the_reference_binary(tmp_path):
    """Linear modifications (overdensity, potential, velocity, angular momentum), splice, reversal, chi2, dump_grid, dump_ps on
    one and two levels in the three draw orders (variance modifications are left to the tool: minutes on the mock device)."""
    rng = random.Random(11)
    compared = 0
    for it in range(24):
        body = fuzz.make_modification_case(rng, allow_variance=False)
        diffs = fuzz.compare_modifications(body, str(tmp_path / ("m%d" % it)))
        if diffs is None:
            continue
        compared += 1
        assert not diffs, (diffs[:4], body)
    assert compared >= 10

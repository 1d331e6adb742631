"""CPU run of tests/test_gpu_grafic.py: the grafic writer, the grafic mapper, the refinement masks and import_level are host
logic (genetic_b200/grafic.py) over a handful of C ABI calls, replaced here by the numpy stand-in tests/host/mock_device.py.
The goldens are the reference's own; the device path runs the same bodies under `-m gpu`."""
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_mock = _load("gx_mock_device", os.path.join(HERE, "host", "mock_device.py"))
_cases = _load("gx_grafic_cases", os.path.join(HERE, "test_gpu_grafic.py"))


@pytest.fixture(autouse=True)
def mock_device(monkeypatch):
    import genetic_b200.icgen as I
    m = _mock.install(monkeypatch)
    monkeypatch.setattr(I, "CAMB", _mock.MockSpectrum)
    return m


def _cpu(fn):
    """The same body without the gpu mark (pytestmark of the source module does not travel with the function)."""
    return fn


test_grafic_output_golden_12 = _cpu(_cases.test_grafic_output_golden_12)
test_grafic_output_and_masks_with_virtual_levels = _cpu(_cases.test_grafic_output_and_masks_with_virtual_levels)
test_refinement_mask_goldens = _cpu(_cases.test_refinement_mask_goldens)
test_grafic_mapper_id_goldens = _cpu(_cases.test_grafic_mapper_id_goldens)
test_import_level_golden_11 = _cpu(_cases.test_import_level_golden_11)
test_centred_grafic_zoom_runs = _cpu(_cases.test_centred_grafic_zoom_runs)

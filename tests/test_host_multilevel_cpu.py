"""CPU tests of the Python HOST LOGIC of the multi-level layer (genetic_b200/multilevel.py): the same test bodies as
tests/test_gpu_multilevel.py, run with the C ABI replaced by a numpy stand-in (tests/host/mock_device.py, test
infrastructure). They check the reference's control flow as restated by the host classes (metric, covectors,
Gram-Schmidt, variance iterations, splice + conjugate gradient) against the oracle's independent restatement and the
reference's goldens; the kernels themselves are checked on the GPU by the `-m gpu` run of the same bodies."""
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
_cases = _load("gx_multilevel_cases", os.path.join(HERE, "test_gpu_multilevel.py"))


@pytest.fixture(autouse=True)
def mock_device(monkeypatch):
    return _mock.install(monkeypatch)


test_chi2_two_levels_golden_02m = _cases.test_chi2_two_levels_golden_02m
test_metric_vs_oracle_three_levels = _cases.test_metric_vs_oracle_three_levels
test_zoom_combination_vs_oracle_and_golden_01c = _cases.test_zoom_combination_vs_oracle_and_golden_01c
test_two_linear_modifications_golden_02c = _cases.test_two_linear_modifications_golden_02c
test_velocity_covector_vs_oracle = _cases.test_velocity_covector_vs_oracle
test_variance_values_golden_02f_02g = _cases.test_variance_values_golden_02f_02g
test_variance_modification_reaches_target_and_matches_oracle = _cases.test_variance_modification_reaches_target_and_matches_oracle
test_splice_golden_24_and_oracle = _cases.test_splice_golden_24_and_oracle
test_overdensity_and_variance_modification_golden_02i = _cases.test_overdensity_and_variance_modification_golden_02i
test_velocity_angular_momentum_potential_modifications_vs_oracle = _cases.test_velocity_angular_momentum_potential_modifications_vs_oracle


_ref_cases = _load("gx_reference_cases", os.path.join(HERE, "test_gpu_vs_reference.py"))
import oracle as _oracle  # noqa: E402
test_zoom_with_modifications_matches_the_reference_binary = pytest.mark.skipif(
    _oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")(_ref_cases.test_zoom_with_modifications_matches_the_reference_binary)
test_zoom_particle_files_match_the_reference_binary = pytest.mark.skipif(
    _oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")(_ref_cases.test_zoom_particle_files_match_the_reference_binary)


_mapper_cases = _load("gx_mapper_cases", os.path.join(HERE, "test_gpu_mapper.py"))
test_gadget_zoom_goldens_10b_10c = _mapper_cases.test_gadget_zoom_goldens_10b_10c
test_example_paramfile_end_to_end_matches_the_reference_binary = pytest.mark.skipif(
    _oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")(_ref_cases.test_example_paramfile_end_to_end_matches_the_reference_binary)


# the reference's own paramfiles through the command layer (genetic_b200/icgen.py), host logic on the mock device
_pf = _load("gx_paramfile_cases", os.path.join(HERE, "test_gpu_paramfile.py"))


@pytest.fixture(autouse=True)
def _mock_spectrum_for_icgen(monkeypatch, mock_device):
    import genetic_b200.icgen as I
    monkeypatch.setattr(I, "CAMB", _mock.MockSpectrum)


for _name in dir(_pf):
    if _name.startswith("test_"):
        globals()["test_paramfile_" + _name[5:]] = getattr(_pf, _name)
test_chunked_evaluation_equals_one_shot = _mapper_cases.test_chunked_evaluation_equals_one_shot


_fg = _load("gx_feature_golden_cases", os.path.join(HERE, "test_gpu_zz_feature_goldens.py"))
test_paramfile_mapper_and_writer_feature_goldens = _fg.test_mapper_and_writer_feature_goldens
test_stray_and_autopad_particle_files_match_the_reference_binary = _fg.test_stray_and_autopad_particle_files_match_the_reference_binary
test_streaming_unigrid_writer_equals_the_mapped_path = _fg.test_streaming_unigrid_writer_equals_the_mapped_path

#!/bin/bash
# Session 23: non-power-of-two virtual grids through the cross-grid gathers (supersample 3), everything that failed in s20b.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_vs_reference.py tests/test_gpu_paramfile.py tests/test_gpu_mapper.py "tests/test_gpu_multilevel.py::test_interpolate_subsample_deinterpolate_vs_oracle" "tests/test_gpu_multilevel.py::test_partial_coarse_cells_and_clamped_parent" -q -m gpu > gpurun_out/s23_pytest.log 2>&1
echo "exit $?" >> gpurun_out/s23_pytest.log
tail -6 gpurun_out/s23_pytest.log

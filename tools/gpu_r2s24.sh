#!/bin/bash
# Round 2 session 24 (1 GPU): config 3 (three-level 256^3 zoom with modifications) with the field-buffer pool and one push per
# iteration; A/B against GX_POOL=0 in the same session; goldens that exercise the same code
mkdir -p gpurun_out
echo "== pool on"; timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing > gpurun_out/r2s24_zoom256_pool.json 2> gpurun_out/r2s24_a.err; tail -3 gpurun_out/r2s24_a.err; cat gpurun_out/r2s24_zoom256_pool.json
echo "== pool off"; GX_POOL=0 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu > gpurun_out/r2s24_zoom256_nopool.json 2> gpurun_out/r2s24_b.err; tail -3 gpurun_out/r2s24_b.err; cat gpurun_out/r2s24_zoom256_nopool.json
( time python -m pytest tests/test_gpu_multilevel.py tests/test_cpp_modif.py tests/test_gpu_zz_feature_goldens.py -q -m gpu ) > gpurun_out/r2s24_pytest.log 2>&1
tail -15 gpurun_out/r2s24_pytest.log

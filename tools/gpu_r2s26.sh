#!/bin/bash
# Round 2 session 26 (1 GPU): config 3 after the grouped gather kernel, table-based filters and the batched window expansion
mkdir -p gpurun_out
echo "== new"; timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing > gpurun_out/r2s26_zoom256.json 2> gpurun_out/r2s26_a.err; tail -3 gpurun_out/r2s26_a.err; cat gpurun_out/r2s26_zoom256.json
echo "== group 1, direct filters"; GX_GATHER_GROUP=1 GX_FILTER_TABLE=0 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing > gpurun_out/r2s26_zoom256_old.json 2> gpurun_out/r2s26_b.err; tail -3 gpurun_out/r2s26_b.err; cat gpurun_out/r2s26_zoom256_old.json
echo "== group 4"; GX_GATHER_GROUP=4 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['seconds'], d['classes_ms']['interp_add'], d['device_ms'])"
echo "== group 16"; GX_GATHER_GROUP=16 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['seconds'], d['classes_ms']['interp_add'], d['device_ms'])"
( time python -m pytest tests/test_gpu_multilevel.py tests/test_cpp_modif.py tests/test_gpu_zz_feature_goldens.py tests/test_gpu_paramfile.py tests/test_gpu_vs_reference.py tests/test_gpu_mapper.py -q -m gpu ) > gpurun_out/r2s26_pytest.log 2>&1
tail -15 gpurun_out/r2s26_pytest.log

#!/bin/bash
# Round 2 session 31 (1 GPU): register-resident gather kernel A/B on config 3, zoom / mapper / grafic tests on it; default
# bench with the pinned e2e buffers allocated on the GPU-local NUMA node
mkdir -p gpurun_out
echo "== small gather"; timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing > gpurun_out/r2s31_zoom256.json 2> gpurun_out/r2s31_a.err; tail -2 gpurun_out/r2s31_a.err
python -c "import json; d=json.loads(open('gpurun_out/r2s31_zoom256.json').read()); print(d['seconds'], d['classes_ms'].get('interp_add'), d['classes_ms'].get('subsample_add'), d['device_ms'], d['sumsq_per_level'])"
echo "== generic gather"; GX_GATHER_SMALL=0 timeout 600 python tools/zoom_bench.py --n 256 --no-cpu --timing 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['seconds'], d['classes_ms'].get('interp_add'), d['classes_ms'].get('subsample_add'), d['device_ms'], d['sumsq_per_level'])"
( time python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_mapper.py tests/test_gpu_zz_feature_goldens.py tests/test_gpu_grafic.py tests/test_gpu_vs_reference.py -q -m gpu ) > gpurun_out/r2s31_pytest.log 2>&1
tail -5 gpurun_out/r2s31_pytest.log
python bench.py > gpurun_out/r2s31_bench2048.json 2> gpurun_out/r2s31_bench.err; tail -2 gpurun_out/r2s31_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r2s31_bench2048.json').readline())
print(round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['ms_per_step'], 1), d['config'].get('e2e_pinned_on_cpus'))"
nvidia-smi topo -m 2>/dev/null | head -12

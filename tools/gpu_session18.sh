# session 18: final state of round 1 -- noise bit-exactness with the adopted residency, slab tests, official bench line
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_noise.py tests/test_gpu_slab.py -m gpu -x -q ) > $O/s18_pytest.log 2>&1
tail -4 $O/s18_pytest.log
timeout 900 python bench.py > $O/s18_bench2048.json 2> $O/s18_a.err
cat $O/s18_bench2048.json

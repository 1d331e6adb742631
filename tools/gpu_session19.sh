# session 19: GPU path against the reference binary (oracle/_ref/genetIC, built in the build container) + reference arm
set -x
O=gpurun_out
mkdir -p $O
ls -la oracle/_ref/genetIC
( time timeout 600 python -m pytest tests/test_gpu_vs_reference.py tests/test_oracle_ref.py -x -q ) > $O/s19_pytest.log 2>&1
tail -6 $O/s19_pytest.log
timeout 300 python bench.py --impl reference > $O/s19_bench_reference.json 2> $O/s19_r.err
cat $O/s19_bench_reference.json

# session 13: column-pass variants at 2048 (8 points per thread, computed twiddle powers)
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python tools/fft_variant_bench.py 2048 --variants=1,6,7,8,9,10,11,12 > $O/s13_cols_variants.txt 2>&1
cat $O/s13_cols_variants.txt

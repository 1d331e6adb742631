#!/bin/sh
# TEST INFRASTRUCTURE. Compiles the UNMODIFIED reference (genetIC) from the sources where they lie under /root/reference
# against the stand-in headers of oracle/shims/ (FFTW3, GSL and HDF5 are not installed in this image). Output:
# oracle/_ref/genetIC (git-ignored, travels to the GPU box with the snapshot). Flags follow genetIC/Makefile:2,23.
# Usage: oracle/build_ref.sh [reference_root]   (default /root/reference)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-/root/reference}/genetIC
[ -f "$REF/src/main.cpp" ] || { echo "reference sources not found under $REF" >&2; exit 3; }
OUT="$HERE/_ref"
mkdir -p "$OUT"
CXX=/usr/bin/g++
CC=/usr/bin/gcc
CODEOPTIONS="-DDOUBLEPRECISION -DOUTPUT_IN_DOUBLEPRECISION -DCUBIC_INTERPOLATION -DFRACTIONAL_K_SPLIT=0.3 -DFRACTIONAL_FILTER_TEMPERATURE=0.1 -DFILTER_ON_COARSE_GRID -DZELDOVICH_GRADIENT_FOURIER_SPACE"
CFLAGS="-O3 -fopenmp -std=c++17 -w -I$HERE/shims -I$REF -DOPENMP -DFFTW3 -DFFTW_THREADS -DGIT_VERSION=\"shim-build\" -DGIT_MODIFIED=\"\""
$CC -O2 -ffp-contract=off -c "$HERE/gx_oracle.c" -o "$OUT/gx_oracle.o"
$CXX -O3 -fopenmp -std=c++17 -I"$HERE/shims" -c "$HERE/shims/fft_shim.cpp" -o "$OUT/fft_shim.o"
$CXX -O2 -std=c++17 -I"$HERE/shims" -c "$HERE/shims/gsl_shim.cpp" -o "$OUT/gsl_shim.o"
for f in main tools/logging tools/progress/progress tools/filesystem; do
  o="$OUT/$(echo $f | tr / _).o"
  $CXX $CFLAGS $CODEOPTIONS -c "$REF/src/$f.cpp" -o "$o" &
done
wait
$CXX -fopenmp -o "$OUT/genetIC" "$OUT"/main.o "$OUT"/tools_logging.o "$OUT"/tools_progress_progress.o "$OUT"/tools_filesystem.o \
  "$OUT/fft_shim.o" "$OUT/gsl_shim.o" "$OUT/gx_oracle.o" -lm -lpthread
echo "$OUT/genetIC"
# The only GPU path the reference ships: USE_CUFFT=1 (genetIC/Makefile:59-70, fourier.hpp:6-8), its FFTs through cuFFTW.
# Same unmodified sources, the real cuFFTW of the CUDA toolkit instead of the FFT stand-in (GSL / HDF5 stand-ins as above).
CUDA=${CUDA_HOME:-/usr/local/cuda}
if [ -f "$CUDA/include/cufftw.h" ] && [ -f "$CUDA/lib64/libcufftw.so" ]; then
  $CXX $CFLAGS $CODEOPTIONS -DUSE_CUFFT -I"$CUDA/include" -c "$REF/src/main.cpp" -o "$OUT/main_cufft.o"
  $CXX -fopenmp -o "$OUT/genetIC_cufft" "$OUT"/main_cufft.o "$OUT"/tools_logging.o "$OUT"/tools_progress_progress.o \
    "$OUT"/tools_filesystem.o "$OUT/gsl_shim.o" "$OUT/gx_oracle.o" -L"$CUDA/lib64" -Wl,-rpath,"$CUDA/lib64" -lcufftw -lcufft -lcudart -lm -lpthread
  echo "$OUT/genetIC_cufft"
fi

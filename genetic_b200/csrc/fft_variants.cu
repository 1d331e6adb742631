// Development entry point: one column pass of the 3-D FFT with an explicit (cluster size, tile width, residency)
// configuration, so tools/fft_variant_bench.py can sweep them on the GPU. Not part of the product path.
#include "../../include/genetic_b200.h"
#include "fft.cuh"

using namespace gx;
using namespace gx::fft;

template <int N, int CL, int TC, int MINB, int SEL = 0> static int run_variant(void *field, int pass, cudaStream_t st) {
  using T = double;
  using C = cplx<T>;
  if (int rc = ensure_twiddles()) return rc;
  C *d = (C *) field;
  int nbatch;
  if (pass == 0) {
    ColGeo g = make_col_geo<PassX>(N, N, N / 2 + 1, nbatch);
    return launch_cols_cfg<T, N, CL, TC, MINB, SEL, 1, PassX>(d, d, g, nbatch, OpIdentity<T>(), OpIdentity<T>(), st);
  }
  ColGeo g = make_col_geo<PassY>(N, N, N / 2 + 1, nbatch);
  return launch_cols_cfg<T, N, CL, TC, MINB, SEL, 1, PassY>(d, d, g, nbatch, OpIdentity<T>(), OpIdentity<T>(), st);
}

extern "C" int gx_dev_fft_cols_variant(void *field, int n, int pass, int variant, void *stream) {
  cudaStream_t st = (cudaStream_t) stream;
  if (n == 2048) {
    switch (variant) {
      case 0: return run_variant<2048, 8, 8, 4, 4>(field, pass, st);
      case 1: return run_variant<2048, 8, 8, 5, 4>(field, pass, st);
      case 2: return run_variant<2048, 8, 8, 6, 4>(field, pass, st);
      case 3: return run_variant<2048, 16, 8, 8, 4>(field, pass, st);
      case 4: return run_variant<2048, 16, 8, 10, 4>(field, pass, st);
      case 5: return run_variant<2048, 16, 16, 4, 4>(field, pass, st);
      case 6: return run_variant<2048, 8, 8, 2, 5>(field, pass, st);    // 8 points per thread, 256-thread CTAs
      case 7: return run_variant<2048, 8, 8, 3, 5>(field, pass, st);
      case 8: return run_variant<2048, 8, 8, 4, 5>(field, pass, st);
      case 9: return run_variant<2048, 8, 8, 5, 12>(field, pass, st);   // twiddle powers by multiplication
      case 10: return run_variant<2048, 8, 8, 4, 12>(field, pass, st);
      case 11: return run_variant<2048, 8, 8, 3, 13>(field, pass, st);  // both
      case 12: return run_variant<2048, 8, 8, 4, 13>(field, pass, st);
    }
  } else if (n == 1024) {
    switch (variant) {
      case 0: return run_variant<1024, 4, 8, 4, 4>(field, pass, st);
      case 1: return run_variant<1024, 4, 8, 5, 4>(field, pass, st);
      case 2: return run_variant<1024, 4, 8, 6, 4>(field, pass, st);
      case 3: return run_variant<1024, 8, 8, 8, 4>(field, pass, st);
      case 4: return run_variant<1024, 8, 8, 10, 4>(field, pass, st);
      case 5: return run_variant<1024, 8, 16, 4, 4>(field, pass, st);
    }
  }
  return gx::fail("gx_dev_fft_cols_variant: unknown variant");
}

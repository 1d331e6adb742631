// Entry points not implemented yet in this build step.
#include "../../include/genetic_b200.h"
#include "gx_common.cuh"
extern "C" {
int gx_zeldovich_offset(const void *delta_k, void *psi, int n, int dtype, double grid_length, int axis, void *stream) {
  int rc = gx_zeldovich_fourier(delta_k, psi, n, dtype, grid_length, axis, stream);
  if (rc) return rc;
  return gx_fft_inverse(psi, n, dtype, stream);
}
}

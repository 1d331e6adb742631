// Entry points not implemented yet in this build step.
#include "../../include/genetic_b200.h"
#include "gx_common.cuh"
extern "C" {
int gx_noise_draw(void *, int, int, unsigned long, int, int, unsigned long long *, void *) { return gx::fail("gx_noise_draw: not implemented yet"); }
int gx_rng_raw(unsigned long, unsigned long long, long, uint32_t *) { return gx::fail("gx_rng_raw: not implemented yet"); }
int gx_rng_gaussians(unsigned long, long, double *, unsigned long long *) { return gx::fail("gx_rng_gaussians: not implemented yet"); }
int gx_zeldovich_offset(const void *delta_k, void *psi, int n, int dtype, double grid_length, int axis, void *stream) {
  int rc = gx_zeldovich_fourier(delta_k, psi, n, dtype, grid_length, axis, stream);
  if (rc) return rc;
  return gx_fft_inverse(psi, n, dtype, stream);
}
}

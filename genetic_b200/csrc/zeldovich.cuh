// Fused Zeldovich passes (internal C++ interface shared by zeldovich.cu and pipeline.cu).
#pragma once
#include "gx_common.cuh"
#include "../../include/genetic_b200.h"

namespace gx {

struct EmitGeo {      // ZeldovichParticleEvaluator / Grid scalars for one axis (zeldovich.hpp:87-119, grid.hpp:519-531,662-669)
  double cell_size, offset_lower, domain_size, vel_ratio;
};

// psi_k(axis) = i k_axis delta_k / k^2 with the Hermitian mirror, inverse-transformed along x and y:
// `work` receives the half-transformed field [x][y][kz]; delta_k is only read.
// prediv != 0: `delta_k` already holds delta_k / k^2 (apply_transfer_ex with divide_k2), the pass only multiplies by i k_axis.
int zeldovich_xy(const void *delta_k, void *work, int n, int dtype, double grid_length, int axis, int prediv,
                 cudaStream_t st);
// Field::applyTransferFunction (field.hpp:394-412) with the covariance evaluated in-kernel, optionally followed by
// the division by k^2 that every Zeldovich component needs
int apply_transfer_ex(void *field, int n, int dtype, double grid_length, const gx_spectrum *s, double power,
                      int divide_k2, cudaStream_t st, int iy0 = 0, int ny = -1);
int build_transfer_table(int n, int dtype, double grid_length, const gx_spectrum *s, double power, int divide_k2,
                         cudaStream_t st, const void **table);
// gx_noise_draw with every mode multiplied by table[kx^2 + ky^2 + kz^2] as it is written (Fourier orders only)
int noise_draw_scaled(void *field, int n, int dtype, unsigned long seed, int mode, int reverse_order,
                      unsigned long long *stream_offset_io, cudaStream_t st, const void *table, void *scratch = nullptr,
                      size_t scratch_bytes = 0);  // scratch: device memory the staged draw may use (e.g. the idle work field)
int unit_variance_slab(void *field, int n, int dtype, int iy0, int ny, cudaStream_t st);
// last pass: c2r along z of rows [row0, row0+nrows) of `work`, scaled by n^-1.5, written back in place (real space)
int zeldovich_z_real(void *work, int n, int dtype, long row0, long nrows, cudaStream_t st);
// last pass fused with the particle evaluation: pos/vel of cells id0 = row0*n ... written to out_pos/out_vel
// (element i = id - id0, unit stride; either pointer may be NULL)
int zeldovich_z_emit(const void *work, int n, int dtype, int axis, const EmitGeo &geo, long row0, long nrows,
                     void *out_pos, void *out_vel, int out_dtype, cudaStream_t st);
// shared-transform schedule, G = delta_k / k^2: G <- x-pass(mirror(G)) in place, then for axis 1 work <- y-pass(i k_y G)
// (out of place) and for axis 2 G <- i k_z y-pass(G) (in place): both ready for the z-pass
int zeldovich_shared_x(void *g, int n, int dtype, cudaStream_t st);
int zeldovich_shared_y(void *g, void *work, int n, int dtype, double grid_length, int axis, cudaStream_t st);

}  // namespace gx

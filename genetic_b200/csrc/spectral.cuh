// Element-wise spectral operators on the padded half-complex layout, and the slot sweep they share.
// Reference: tools/numerics/fourier.hpp:110-151 (mode iteration), simulation/field/field.hpp:394-412,
// 741-746, cosmology/camb.hpp:101-139,204-216, simulation/particles/zeldovich.hpp:152-181.
#pragma once
#include "gx_common.cuh"

namespace gx {

// Device view of a power spectrum (camb.hpp:173-216 / :150-168)
struct SpecDev {
  int kind;           // 0 = CAMB table, 1 = power law
  int m;              // number of spline knots
  const double *tab;  // device: lx[m], ly[m], c[m]
  double amplitude, ns;
  float amplitude_f, ns_f;
};

// IEEE operations without FMA contraction, so that element-wise results match the reference's
// (x86-64, no -march flags => no FMA) bit for bit.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }

// Grid::wrapIndividualCoordinate (grid.hpp:527-531): x = fmod(x, L); if (x < 0) x += L.  fmod is exact, and for
// L <= x < 2L it equals x - L, which is itself exact (Sterbenz), so the common cases never enter the slow routine.
template <typename T> __device__ __forceinline__ T wrap_coordinate(T x, T L) {
  // -L < x < 2L (every particle: |psi| << L): selects only, no divergent branches; anything else takes the library fmod
  T y = (x >= L) ? x - L : x;
  if (!(x > -L && x < L + L)) y = fmod(x, L);
  return (y < 0) ? add_rn(y, L) : y;
}

// k^2 and |k| of a storage slot exactly as the reference forms them:
//   kx = int * kMin (CoordinateType), k2 = kx*kx + ky*ky + kz*kz, k = sqrt(k2)
template <typename T> __device__ __forceinline__ void slot_k(int n, int ix, int iy, int iz, T kmin, T &kx, T &ky, T &kz,
                                                             T &k2) {
  kx = mul_rn((T) signed_mode(ix, n), kmin);
  ky = mul_rn((T) signed_mode(iy, n), kmin);
  kz = mul_rn((T) iz, kmin);
  k2 = add_rn(add_rn(mul_rn(kx, kx), mul_rn(ky, ky)), mul_rn(kz, kz));
}

// natural cubic spline evaluation, GSL cspline_eval form (interpolation.hpp:72-76 -> gsl_spline_eval)
__device__ __forceinline__ double spline_eval(const double *lx, const double *ly, const double *c, int m, double x) {
  int lo = 0, hi = m - 1;
  while (hi > lo + 1) {
    int mid = (hi + lo) >> 1;
    if (lx[mid] > x) hi = mid; else lo = mid;
  }
  const double dx = lx[lo + 1] - lx[lo];
  const double dy = ly[lo + 1] - ly[lo];
  const double delx = x - lx[lo];
  const double c_i = c[lo], c_ip1 = c[lo + 1];
  const double b_i = __dsub_rn(__ddiv_rn(dy, dx), __ddiv_rn(__dmul_rn(dx, __dadd_rn(c_ip1, __dmul_rn(2.0, c_i))), 3.0));
  const double d_i = __ddiv_rn(__dsub_rn(c_ip1, c_i), __dmul_rn(3.0, dx));
  return __dadd_rn(ly[lo], __dmul_rn(delx, __dadd_rn(b_i, __dmul_rn(delx, __dadd_rn(c_i, __dmul_rn(delx, d_i))))));
}

// CAMB::operator()(k) / PowerLawPowerSpectrum::operator()(k) in the field's CoordinateType
template <typename T>
__device__ __forceinline__ T spectrum_power(const SpecDev &sp, const double *lx, const double *ly, const double *c, T k);

template <>
__device__ __forceinline__ double spectrum_power<double>(const SpecDev &sp, const double *lx, const double *ly,
                                                         const double *c, double k) {
  if (sp.kind == 1) return k == 0 ? 0.0 : sp.amplitude * pow(k, sp.ns);
  double lt = 0.0;
  if (k != 0) {
    const double lg = log(k);
    if (!(lg < lx[0] || lg > lx[sp.m - 1])) lt = exp(spline_eval(lx, ly, c, sp.m, lg));
  }
  return __dmul_rn(__dmul_rn(__dmul_rn(sp.amplitude, pow(k, sp.ns)), lt), lt);
}

template <>
__device__ __forceinline__ float spectrum_power<float>(const SpecDev &sp, const double *lx, const double *ly,
                                                       const double *c, float k) {
  if (sp.kind == 1) return k == 0 ? 0.0f : sp.amplitude_f * powf(k, sp.ns_f);
  float lt = 0.0f;
  if (k != 0) {
    const double lg = log((double) k);
    if (!(lg < lx[0] || lg > lx[sp.m - 1])) lt = (float) exp(spline_eval(lx, ly, c, sp.m, lg));
  }
  return __fmul_rn(__fmul_rn(__fmul_rn(sp.amplitude_f, powf(k, sp.ns_f)), lt), lt);
}

// spec^power as Field::applyTransferFunction forms it (field.hpp:405-406)
template <typename T> __device__ __forceinline__ T spec_pow(T spec, double power) {
  if (power != 1.0 && spec != (T) 0) {
    if (power == 0.5) return sqrt_rn(spec);  // == pow(x, 0.5) wherever pow is correctly rounded
    return (T) pow((double) spec, power);
  }
  return spec;
}

// ---- slot sweep: grid.y = ix, grid.x * blockDim.x covers the ny*nc slots of one x-plane ---------------
// The functor sees the global storage indices (ix, iy, iz) and the element offset s in the array it was given: the
// whole field [n][n][nc] (iy0 = 0, ny = n) or one rank's ky-slab [n][ny][nc] holding iy in [iy0, iy0 + ny).
template <class F> __global__ void __launch_bounds__(256) slot_sweep_kernel(int n, int nc, int iy0, int ny, F f) {
  const int ix = blockIdx.y;
  const int plane = ny * nc;
  f.init();  // optional per-CTA prologue (e.g. stage a table in shared memory); must be called by all threads
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int iyl = i / nc;
    const int iz = i - iyl * nc;
    f(ix, iy0 + iyl, iz, (size_t) ix * plane + i);
  }
}

template <class F>
int launch_slot_sweep(int n, F f, cudaStream_t st, size_t smem = 0, const char *name = "slot_sweep", int iy0 = 0,
                      int ny = -1) {
  GX_TIMED(name, st);
  const int nc = n / 2 + 1;
  if (ny < 0) ny = n;
  const int plane = ny * nc;
  int bx = (plane + 255) / 256;
  if (bx > 64) bx = 64;  // 64 x n CTAs: >= 4 waves of 148 SMs from n = 16 upward
  dim3 grid(bx, n);
  // a CAMB table with more than ~2000 usable rows needs more than the default 48 KB of dynamic shared memory for its spline
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return fail("transfer-function table too long for the in-kernel spline (more than ~8500 rows)");
    GX_CUDA_CHECK(cudaFuncSetAttribute(slot_sweep_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  }
  slot_sweep_kernel<F><<<grid, 256, smem, st>>>(n, nc, iy0, ny, f);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

}  // namespace gx

// Reductions and BLAS-1 style sweeps over fields (warp-shuffle + block reduction, fixed-order final sum).
// Reference: MultiLevelField::innerProduct (multilevelfield.hpp:248-294) with the mode weights of
// fourier.hpp:116-147; Field::innerProduct / addScaled / operator*= (field.hpp:306-377).
#include "../../include/genetic_b200.h"
#include "gx_common.cuh"

namespace gx {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block partial -> partials[blockIdx]; deterministic (no atomics), final sum on a single CTA
__device__ __forceinline__ void block_store_sum(double v, double *partials, int block_linear) {
  __shared__ double s[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < (blockDim.x >> 5)) ? s[l] : 0.0;
    v = warp_sum(v);
    if (l == 0) partials[block_linear] = v;
  }
}

__global__ void final_sum_kernel(const double *partials, int n, double *out) {
  double v = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  __shared__ double s[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < (blockDim.x >> 5)) ? s[l] : 0.0;
    v = warp_sum(v);
    if (l == 0) *out = v;
  }
}

// weight of slot (ix,iy,iz) in the reference's half-space accumulation (0 for slots it never visits)
__device__ __forceinline__ double slot_weight(int n, int ix, int iy, int iz) {
  if (!slot_is_independent(n, ix, iy, iz)) return 0.0;
  const int h = n >> 1;
  const bool sx = (ix == 0 || ix == h), sy = (iy == 0 || iy == h), sz = (iz == 0 || iz == h);
  return (sx && sy && sz) ? 1.0 : 2.0;
}

template <typename T> __global__ void __launch_bounds__(256)
dot_fourier_kernel(const cplx<T> *__restrict__ a, const cplx<T> *__restrict__ b, int n, int nc, double *partials) {
  const int ix = blockIdx.y;
  const int plane = n * nc;
  double acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int iy = i / nc, iz = i - iy * nc;
    const double w = slot_weight(n, ix, iy, iz);
    const size_t s = (size_t) ix * plane + i;
    const cplx<T> u = a[s], v = b[s];
    acc += w * ((double) u.x * (double) v.x + (double) u.y * (double) v.y);
  }
  block_store_sum(acc, partials, blockIdx.y * gridDim.x + blockIdx.x);
}

template <typename T> __global__ void __launch_bounds__(256)
dot_real_kernel(const T *__restrict__ a, const T *__restrict__ b, int n, double *partials) {
  // rows of n reals at pitch n+2
  const long nrows = (long) n * n;
  double acc = 0;
  for (long r = blockIdx.x; r < nrows; r += gridDim.x) {
    const T *ra = a + r * (n + 2), *rb = b + r * (n + 2);
    for (int z = threadIdx.x; z < n; z += blockDim.x) acc += (double) ra[z] * (double) rb[z];
  }
  block_store_sum(acc, partials, blockIdx.x);
}

template <typename T> __global__ void scale_kernel(T *f, size_t len, T a) {
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t) gridDim.x * blockDim.x)
    f[i] = f[i] * a;
}

template <typename T> __global__ void axpy_kernel(T *y, const T *x, size_t len, T a) {
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t) gridDim.x * blockDim.x)
    y[i] = y[i] + a * x[i];  // data[i] += scale*other[i]; contraction to FMA is within the reductions' tolerance
}

template <typename T> __global__ void __launch_bounds__(256)
axpy_fourier_kernel(cplx<T> *y, const cplx<T> *x, int n, int nc, T a) {
  const int ix = blockIdx.y;
  const int plane = n * nc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int iy = i / nc, iz = i - iy * nc;
    if (!slot_is_independent(n, ix, iy, iz)) continue;
    const size_t s = (size_t) ix * plane + i;
    cplx<T> u = y[s];
    const cplx<T> v = x[s];
    u.x += a * v.x; u.y += a * v.y;
    y[s] = u;
  }
}

template <typename T> __global__ void mul_mask_kernel(T *f, const unsigned char *mask, int n) {
  const long nrows = (long) n * n;
  for (long r = blockIdx.x; r < nrows; r += gridDim.x) {
    T *row = f + r * (n + 2);
    const unsigned char *m = mask + r * n;
    for (int z = threadIdx.x; z < n; z += blockDim.x) row[z] = row[z] * (T) (signed char) m[z];
  }
}

static double *g_partials = nullptr;  // 64*4096 partials + 1 result
static const int kMaxPartials = 64 * 4096;
static int ensure_partials() {
  if (g_partials) return 0;
  GX_CUDA_CHECK(cudaMalloc(&g_partials, (kMaxPartials + 1) * sizeof(double)));
  return 0;
}
void release_reduce_workspace() {  // gx_workspace_release / gx_set_device(another device)
  if (g_partials) { cudaDeviceSynchronize(); cudaFree(g_partials); }
  g_partials = nullptr;
}
static int check_n_reduce(int n) {  // the partials hold 64 blocks x n planes: n <= 4096
  if (n < 4 || n > 4096 || (n & 1)) return fail("grid size must be even and in [4,4096]");
  return 0;
}

template <typename T> static int dot_fourier_t(const void *a, const void *b, int n, double *host_out, cudaStream_t st) {
  if (check_n_reduce(n) || ensure_partials()) return 1;
  const int nc = n / 2 + 1, plane = n * nc;
  int bx = (plane + 255) / 256; if (bx > 64) bx = 64;
  dot_fourier_kernel<T><<<dim3(bx, n), 256, 0, st>>>((const cplx<T> *) a, (const cplx<T> *) b, n, nc, g_partials);
  final_sum_kernel<<<1, 1024, 0, st>>>(g_partials, bx * n, g_partials + kMaxPartials);
  count_launch(2);
  GX_LAUNCH_CHECK();
  GX_CUDA_CHECK(cudaMemcpyAsync(host_out, g_partials + kMaxPartials, sizeof(double), cudaMemcpyDeviceToHost, st));
  GX_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

template <typename T> static int dot_real_t(const void *a, const void *b, int n, double *host_out, cudaStream_t st) {
  if (check_n_reduce(n) || ensure_partials()) return 1;
  long nrows = (long) n * n;
  int grid = (int) (nrows < 148 * 16 ? nrows : 148 * 16);
  dot_real_kernel<T><<<grid, 256, 0, st>>>((const T *) a, (const T *) b, n, g_partials);
  final_sum_kernel<<<1, 1024, 0, st>>>(g_partials, grid, g_partials + kMaxPartials);
  count_launch(2);
  GX_LAUNCH_CHECK();
  GX_CUDA_CHECK(cudaMemcpyAsync(host_out, g_partials + kMaxPartials, sizeof(double), cudaMemcpyDeviceToHost, st));
  GX_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace gx

using namespace gx;

#define GX_DISPATCH(dtype, CALL)                              \
  do {                                                        \
    if ((dtype) == GX_F64) { using T = double; return CALL; } \
    if ((dtype) == GX_F32) { using T = float; return CALL; }  \
    return gx::fail("bad dtype");                             \
  } while (0)

template <typename T> static int scale_t(void *f, int n, double a, cudaStream_t st) {
  if (n < 1 || n > 4096) return gx::fail("grid size out of range");
  const size_t len = (size_t) 2 * n * n * (n / 2 + 1);
  scale_kernel<T><<<148 * 8, 256, 0, st>>>((T *) f, len, (T) a);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}
template <typename T> static int axpy_t(void *y, const void *x, int n, double a, cudaStream_t st) {
  const size_t len = (size_t) 2 * n * n * (n / 2 + 1);
  axpy_kernel<T><<<148 * 8, 256, 0, st>>>((T *) y, (const T *) x, len, (T) a);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}
template <typename T> static int axpy_fourier_t(void *y, const void *x, int n, double a, cudaStream_t st) {
  const int nc = n / 2 + 1, plane = n * nc;
  int bx = (plane + 255) / 256; if (bx > 64) bx = 64;
  axpy_fourier_kernel<T><<<dim3(bx, n), 256, 0, st>>>((cplx<T> *) y, (const cplx<T> *) x, n, nc, (T) a);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}
template <typename T> static int mul_mask_t(void *f, const unsigned char *mask, int n, cudaStream_t st) {
  long nrows = (long) n * n;
  int grid = (int) (nrows < 148 * 16 ? nrows : 148 * 16);
  mul_mask_kernel<T><<<grid, 256, 0, st>>>((T *) f, mask, n);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

extern "C" {
int gx_scale(void *field, int n, int dtype, double a, void *stream) { GX_DISPATCH(dtype, scale_t<T>(field, n, a, (cudaStream_t) stream)); }
int gx_axpy(void *y, const void *x, int n, int dtype, double a, void *stream) { GX_DISPATCH(dtype, axpy_t<T>(y, x, n, a, (cudaStream_t) stream)); }
int gx_axpy_fourier(void *y, const void *x, int n, int dtype, double a, void *stream) { GX_DISPATCH(dtype, axpy_fourier_t<T>(y, x, n, a, (cudaStream_t) stream)); }
int gx_mul_mask(void *field, const unsigned char *mask, int n, int dtype, void *stream) { GX_DISPATCH(dtype, mul_mask_t<T>(field, mask, n, (cudaStream_t) stream)); }
int gx_dot_fourier(const void *a, const void *b, int n, int dtype, double *host_out, void *stream) { GX_DISPATCH(dtype, dot_fourier_t<T>(a, b, n, host_out, (cudaStream_t) stream)); }
int gx_dot_real(const void *a, const void *b, int n, int dtype, double *host_out, void *stream) { GX_DISPATCH(dtype, dot_real_t<T>(a, b, n, host_out, (cudaStream_t) stream)); }

int gx_get_mode(const void *field, int n, int dtype, int kx, int ky, int kz, double *re, double *im) {
  // getRealCoeffLocationAndConjugation (fourier.hpp:300-319)
  bool conj = false;
  if (kz < 0) { conj = true; kx = -kx; ky = -ky; kz = -kz; }
  if (ky < 0) ky += n;
  if (kx < 0) kx += n;
  const size_t slot = (size_t) kz + (size_t) (n / 2 + 1) * ((size_t) ky + (size_t) n * kx);
  if (dtype == GX_F64) {
    double v[2];
    GX_CUDA_CHECK(cudaMemcpy(v, (const double *) field + 2 * slot, 16, cudaMemcpyDeviceToHost));
    *re = v[0]; *im = conj ? -v[1] : v[1];
  } else {
    float v[2];
    GX_CUDA_CHECK(cudaMemcpy(v, (const float *) field + 2 * slot, 8, cudaMemcpyDeviceToHost));
    *re = v[0]; *im = conj ? -v[1] : v[1];
  }
  return 0;
}

int gx_set_mode(void *field, int n, int dtype, int kx, int ky, int kz, double re, double im) {
  bool conj = false;
  if (kz < 0) { conj = true; kx = -kx; ky = -ky; kz = -kz; }
  if (ky < 0) ky += n;
  if (kx < 0) kx += n;
  const size_t slot = (size_t) kz + (size_t) (n / 2 + 1) * ((size_t) ky + (size_t) n * kx);
  if (conj) im = -im;
  if (dtype == GX_F64) {
    double v[2] = {re, im};
    GX_CUDA_CHECK(cudaMemcpy((double *) field + 2 * slot, v, 16, cudaMemcpyHostToDevice));
  } else {
    float v[2] = {(float) re, (float) im};
    GX_CUDA_CHECK(cudaMemcpy((float *) field + 2 * slot, v, 8, cudaMemcpyHostToDevice));
  }
  return 0;
}
}  // extern "C"

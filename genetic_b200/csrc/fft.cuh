// Pass kernels of the 3-D FFT (templates; instantiated from fft.cu and the fused Zeldovich path).
#pragma once
#include <vector>
#include <cmath>
#include <type_traits>
#include "gx_common.cuh"
#include "fft_core.cuh"

namespace gx {
namespace fft {

int ensure_twiddles();
template <typename T> const cplx<T> *twiddle_table(int n);
template <typename T> int fft3d(T *data, int n, int dir, cudaStream_t st);

// ---- element operators fused into the load or store side of a column pass -----------------------
template <typename T> struct OpIdentity {
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int) const { return v; }
};
template <typename T> struct OpScale {
  T s;
  explicit OpScale(T s_) : s(s_) {}
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int) const { return mk<T>(v.x * s, v.y * s); }
};

// ---- which axis a column pass transforms ---------------------------------------------------------
// PassY: batch = ix, line index = iy, column = iz            (line stride nc,   batch stride n*nc)
// PassX: one batch, line index = ix, column = iy*nc + iz     (line stride n*nc)
struct PassY {};
struct PassX {};

template <typename T, int N> struct ColsCfg {
  using B = BlockFFT<T, N, -1>;
  static constexpr int P = B::P;
  static constexpr int minTC = (sizeof(T) == 8) ? (N >= 2048 ? 2 : 4) : (N >= 2048 ? 4 : 8);
  static constexpr int byThreads = 256 / P;
  static constexpr int TC = byThreads > minTC ? byThreads : minTC;
  static constexpr int threads = TC * P;
  static constexpr int pitch = line_pitch<T>(N);
  static constexpr size_t smem = (size_t) TC * pitch * sizeof(cplx<T>);
};

// One CTA transforms TC adjacent columns (contiguous in memory) of length N.
template <typename T, int N, int DIR, class Pass, class LoadOp, class StoreOp>
__global__ void __launch_bounds__(ColsCfg<T, N>::threads)
fft_cols_kernel(const cplx<T> *__restrict__ src, cplx<T> *__restrict__ dst, int n, int nc, long ncols,
                long line_stride, long batch_stride, int tiles_per_batch, const cplx<T> *__restrict__ tw,
                LoadOp lop, StoreOp sop) {
  using C = cplx<T>;
  using B = BlockFFT<T, N, DIR>;
  using Cfg = ColsCfg<T, N>;
  constexpr int TC = Cfg::TC, E = B::E, P = B::P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);

  const int c = threadIdx.x % TC;
  const int t = threadIdx.x / TC;
  const long tile = blockIdx.x;
  const int batch = (int) (tile / tiles_per_batch);
  const long col = (tile % tiles_per_batch) * TC + c;
  const bool active = col < ncols;
  int cx, cy, cz;  // slot coordinates that do not depend on the line index
  if (std::is_same<Pass, PassY>::value) { cx = batch; cy = 0; cz = (int) col; }
  else { cx = 0; cy = (int) (col / nc); cz = (int) (col % nc); }

  const long base = (long) batch * batch_stride + col;
  C v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int row = t + i * P;
    if (active) {
      C x = src[base + (long) row * line_stride];
      v[i] = std::is_same<Pass, PassY>::value ? lop(x, cx, row, cz) : lop(x, row, cy, cz);
    } else {
      v[i] = mk<T>(0, 0);
    }
  }
  C *line = smem + c * Cfg::pitch;
  B::run(v, t, line, tw, [&](int k, C val) {
    if (active) {
      C y = std::is_same<Pass, PassY>::value ? sop(val, cx, k, cz) : sop(val, k, cy, cz);
      dst[base + (long) k * line_stride] = y;
    }
  });
}

template <typename T, int N, int DIR, class Pass, class LoadOp, class StoreOp>
int launch_cols(const cplx<T> *src, cplx<T> *dst, int n, LoadOp lop, StoreOp sop, cudaStream_t st) {
  using Cfg = ColsCfg<T, N>;
  const int nc = n / 2 + 1;
  long ncols, line_stride, batch_stride;
  int nbatch;
  if (std::is_same<Pass, PassY>::value) { ncols = nc; line_stride = nc; batch_stride = (long) n * nc; nbatch = n; }
  else { ncols = (long) n * nc; line_stride = (long) n * nc; batch_stride = 0; nbatch = 1; }
  const int tiles_per_batch = (int) ((ncols + Cfg::TC - 1) / Cfg::TC);
  const long grid = (long) tiles_per_batch * nbatch;
  auto kern = fft_cols_kernel<T, N, DIR, Pass, LoadOp, StoreOp>;
  if (Cfg::smem > 48 * 1024) {
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
  }
  const char *tname = std::is_same<Pass, PassY>::value ? "fft_cols_y" : "fft_cols_x";
  GX_TIMED(tname, st);
  kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(src, dst, n, nc, ncols, line_stride, batch_stride,
                                                        tiles_per_batch, twiddle_table<T>(N), lop, sop);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

// ---- row passes (z axis, contiguous): r2c and c2r through a half-length complex FFT ---------------
template <typename T, int N> struct RowsCfg {
  static constexpr int M = N / 2;
  using B = BlockFFT<T, M, -1>;
  static constexpr int P = B::P;
  static constexpr int pitch = line_pitch<T>(M);
  static constexpr int zpitch = M + 2;
  static constexpr size_t per_row = (size_t) (pitch + zpitch) * sizeof(cplx<T>);
  static constexpr int byThreads = (256 / P) > 0 ? (256 / P) : 1;
  static constexpr int bySmem = (int) ((96 * 1024) / per_row) > 0 ? (int) ((96 * 1024) / per_row) : 1;
  static constexpr int LR = byThreads < bySmem ? byThreads : bySmem;
  static constexpr int threads = LR * P;
  static constexpr size_t smem = (size_t) LR * per_row;
};

// Forward: each row of N reals (viewed as M = N/2 complex z[m] = x[2m] + i x[2m+1]) -> N/2+1 complex, in place.
//   X[k] = 1/2 (Z[k] + conj Z[M-k]) - i/2 W_N^k (Z[k] - conj Z[M-k]),  k = 0..M,  Z[M] == Z[0]
template <typename T, int N>
__global__ void __launch_bounds__(RowsCfg<T, N>::threads)
fft_rows_r2c_kernel(T *__restrict__ data, long nrows, const cplx<T> *__restrict__ twM,
                    const cplx<T> *__restrict__ twN) {
  using C = cplx<T>;
  using Cfg = RowsCfg<T, N>;
  using B = BlockFFT<T, Cfg::M, -1>;
  constexpr int M = Cfg::M, E = B::E, P = B::P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  const int l = threadIdx.x / P, t = threadIdx.x % P;
  const long row = (long) blockIdx.x * Cfg::LR + l;
  const bool active = row < nrows;
  C *rowp = reinterpret_cast<C *>(data + row * (long) (N + 2));
  C *line = smem + (size_t) l * (Cfg::pitch + Cfg::zpitch);
  C *zbuf = line + Cfg::pitch;
  C v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) v[i] = active ? rowp[t + i * P] : mk<T>(0, 0);
  B::run(v, t, line, twM, [&](int k, C val) { zbuf[k] = val; });
  __syncthreads();
  if (!active) return;
  const T half = (T) 0.5;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int k = t + i * P;
    const C zk = zbuf[k];
    const C zm = cconj(zbuf[(M - k) & (M - 1)]);
    const C xe = mk<T>((zk.x + zm.x) * half, (zk.y + zm.y) * half);
    const C d = csub(zk, zm);
    const C xo = mk<T>(d.y * half, -d.x * half);  // -i/2 * d
    const C w = __ldg(twN + k);
    rowp[k] = cadd(xe, cmul(w, xo));
    if (k == 0) rowp[M] = mk<T>(zk.x - zk.y, (T) 0);  // X[M] = Re Z0 - Im Z0
  }
}

// Inverse: N/2+1 complex -> N reals in place (unnormalised c2r, times `scale`).
//   Zt[k] = (X[k] + conj X[M-k]) + i W_N^{-k} (X[k] - conj X[M-k]),  z = IFFT_M(Zt),  x[2m] = Re z[m], x[2m+1] = Im z[m]
// The imaginary parts of X[0] and X[M] are ignored, as FFTW's c2r does.
template <typename T, int N>
__global__ void __launch_bounds__(RowsCfg<T, N>::threads)
fft_rows_c2r_kernel(T *__restrict__ data, long nrows, T scale, const cplx<T> *__restrict__ twM,
                    const cplx<T> *__restrict__ twN) {
  using C = cplx<T>;
  using Cfg = RowsCfg<T, N>;
  using B = BlockFFT<T, Cfg::M, 1>;
  constexpr int M = Cfg::M, E = B::E, P = B::P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  const int l = threadIdx.x / P, t = threadIdx.x % P;
  const long row = (long) blockIdx.x * Cfg::LR + l;
  const bool active = row < nrows;
  C *rowp = reinterpret_cast<C *>(data + row * (long) (N + 2));
  C *line = smem + (size_t) l * (Cfg::pitch + Cfg::zpitch);
  C *zbuf = line + Cfg::pitch;
#pragma unroll
  for (int i = 0; i < E; ++i) zbuf[t + i * P] = active ? rowp[t + i * P] : mk<T>(0, 0);
  if (t == 0) zbuf[M] = active ? rowp[M] : mk<T>(0, 0);
  __syncthreads();
  C v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int k = t + i * P;
    C xk = zbuf[k];
    C xm = cconj(zbuf[M - k]);
    if (k == 0) { xk.y = 0; xm.y = 0; }
    const C s = cadd(xk, xm);
    const C d = csub(xk, xm);
    const C w = cconj(__ldg(twN + k));
    const C wd = cmul(w, d);
    v[i] = mk<T>(s.x - wd.y, s.y + wd.x);  // s + i*wd
  }
  __syncthreads();  // everyone has consumed zbuf before it is reused for the outputs
  B::run(v, t, line, twM, [&](int k, C val) { zbuf[k] = val; });
  __syncthreads();
  if (!active) return;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const C z = zbuf[t + i * P];
    rowp[t + i * P] = mk<T>(z.x * scale, z.y * scale);
  }
}

template <typename T, int N> int launch_rows_r2c(T *data, long nrows, cudaStream_t st) {
  using Cfg = RowsCfg<T, N>;
  auto kern = fft_rows_r2c_kernel<T, N>;
  if (Cfg::smem > 48 * 1024)
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
  const long grid = (nrows + Cfg::LR - 1) / Cfg::LR;
  GX_TIMED("fft_rows_r2c", st);
  kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, nrows, twiddle_table<T>(N / 2), twiddle_table<T>(N));
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int N> int launch_rows_c2r(T *data, long nrows, T scale, cudaStream_t st) {
  using Cfg = RowsCfg<T, N>;
  auto kern = fft_rows_c2r_kernel<T, N>;
  if (Cfg::smem > 48 * 1024)
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
  const long grid = (nrows + Cfg::LR - 1) / Cfg::LR;
  GX_TIMED("fft_rows_c2r", st);
  kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, nrows, scale, twiddle_table<T>(N / 2),
                                                        twiddle_table<T>(N));
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

}  // namespace fft
}  // namespace gx

// 3-D real transform for even grid sizes that are not powers of two: three passes of the mixed-radix line transform of
// fft_generic_core.cuh (z: r2c / c2r through a full complex line, then y and x), each line tile staged in shared memory.
// Same storage, same unitary scaling (1 / sqrt(N) per pass) and same mirror-before-inverse contract as the power-of-two
// passes of fft.cuh (FieldFourierManager::performTransform, fourier.hpp:479-518). The power-of-two sizes never come here.
#include "fft_generic_core.cuh"
#include <cmath>
#include <mutex>
#include <vector>

namespace gx {
namespace fftgen {

enum { AX_X = 0, AX_Y = 1, AX_Z_R2C = 2, AX_Z_C2R = 3 };

template <typename T, int AX>
__global__ void __launch_bounds__(256) lines_kernel(T *data, Plan plan, const C2<T> *__restrict__ tw, int dir, int tc, T scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = plan.n, nc = n / 2 + 1;
  C2<T> *in = reinterpret_cast<C2<T> *>(smem_raw);
  C2<T> *out = in + (size_t) n * tc;
  C2<T> *cdata = reinterpret_cast<C2<T> *>(data);
  const long n_lines = (AX == AX_X || AX == AX_Y) ? (long) n * nc : (long) n * n;
  const long l0 = (long) blockIdx.x * tc;
  const int cnt = (int) (n_lines - l0 < tc ? n_lines - l0 : tc);
  auto column_base = [&](long l) -> size_t {  // first element of complex line l of a column pass
    if (AX == AX_X) return (size_t) l;
    return (size_t) (l / nc) * n * nc + (size_t) (l % nc);
  };
  const size_t cstride = AX == AX_X ? (size_t) n * nc : (size_t) nc;
  if (AX == AX_X || AX == AX_Y) {
    for (int e = threadIdx.x; e < n * cnt; e += blockDim.x) {
      const int idx = e / cnt, c = e - idx * cnt;
      in[(size_t) idx * tc + c] = cdata[column_base(l0 + c) + (size_t) idx * cstride];
    }
  } else if (AX == AX_Z_R2C) {
    for (int e = threadIdx.x; e < n * cnt; e += blockDim.x) {
      const int c = e / n, idx = e - c * n;
      in[(size_t) idx * tc + c] = C2<T>{data[(size_t) (l0 + c) * 2 * nc + idx], (T) 0};
    }
  } else {
    for (int e = threadIdx.x; e < nc * cnt; e += blockDim.x) {  // Hermitian completion of the stored half line
      const int c = e / nc, k = e - c * nc;
      const C2<T> v = cdata[(size_t) (l0 + c) * nc + k];
      in[(size_t) k * tc + c] = v;
      if (k != 0 && k != n - k) in[(size_t) (n - k) * tc + c] = C2<T>{v.x, -v.y};
    }
  }
  __syncthreads();
  int Ns = 1;
  for (int f = 0; f < plan.nf; ++f) {
    const int R = plan.radix[f];
    const int items = (n / R) * cnt;
    for (int w = threadIdx.x; w < items; w += blockDim.x) {
      const int j = w / cnt, c = w - j * cnt;
      stage_item<T>(in, out, n, tc, R, Ns, tw, dir, j, c);
    }
    __syncthreads();
    Ns *= R;
    C2<T> *t = in; in = out; out = t;
  }
  if (AX == AX_X || AX == AX_Y) {
    for (int e = threadIdx.x; e < n * cnt; e += blockDim.x) {
      const int idx = e / cnt, c = e - idx * cnt;
      const C2<T> v = in[(size_t) idx * tc + c];
      cdata[column_base(l0 + c) + (size_t) idx * cstride] = C2<T>{v.x * scale, v.y * scale};
    }
  } else if (AX == AX_Z_R2C) {
    for (int e = threadIdx.x; e < nc * cnt; e += blockDim.x) {
      const int c = e / nc, k = e - c * nc;
      const C2<T> v = in[(size_t) k * tc + c];
      cdata[(size_t) (l0 + c) * nc + k] = C2<T>{v.x * scale, v.y * scale};
    }
  } else {
    for (int e = threadIdx.x; e < n * cnt; e += blockDim.x) {
      const int c = e / n, idx = e - c * n;
      data[(size_t) (l0 + c) * 2 * nc + idx] = in[(size_t) idx * tc + c].x * scale;
    }
  }
}

// twiddle tables, one per (n, element type, device); dropped by gx_workspace_release
struct TwEntry { int n, esize, device; void *dev; };
static std::vector<TwEntry> g_tw;
static std::mutex g_tw_mu;

template <typename T> static int twiddles(int n, const C2<T> **out) {
  std::lock_guard<std::mutex> lock(g_tw_mu);
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  for (auto &e : g_tw)
    if (e.n == n && e.esize == (int) sizeof(T) && e.device == dev) { *out = (const C2<T> *) e.dev; return 0; }
  std::vector<C2<T>> h(n);
  for (int t = 0; t < n; ++t) {
    const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) n;
    h[t] = C2<T>{(T) cosl(ang), (T) sinl(ang)};
  }
  void *d = nullptr;
  GX_CUDA_CHECK(cudaMalloc(&d, sizeof(C2<T>) * n));
  GX_CUDA_CHECK(cudaMemcpy(d, h.data(), sizeof(C2<T>) * n, cudaMemcpyHostToDevice));
  g_tw.push_back(TwEntry{n, (int) sizeof(T), dev, d});
  *out = (const C2<T> *) d;
  return 0;
}

void release_workspace() {
  std::lock_guard<std::mutex> lock(g_tw_mu);
  if (g_tw.empty()) return;
  cudaDeviceSynchronize();
  for (auto &e : g_tw) cudaFree(e.dev);
  g_tw.clear();
}

template <typename T, int AX> static int run_pass(T *field, const Plan &plan, const C2<T> *tw, int dir, cudaStream_t st, const char *name) {
  const int n = plan.n, nc = n / 2 + 1;
  const size_t per_line = 2 * (size_t) n * sizeof(C2<T>);
  int tc = (int) ((size_t) 160 * 1024 / per_line);
  if (tc < 1) return fail("grid size too large for the mixed-radix transform (the two line buffers must fit in shared memory)");
  if (tc > 8) tc = 8;
  const size_t smem = per_line * tc;
  const long n_lines = (AX == AX_X || AX == AX_Y) ? (long) n * nc : (long) n * n;
  GX_TIMED(name, st);
  if (smem > 48 * 1024)
    GX_CUDA_CHECK(cudaFuncSetAttribute(lines_kernel<T, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  lines_kernel<T, AX><<<(unsigned) ((n_lines + tc - 1) / tc), 256, smem, st>>>(field, plan, tw, dir, tc, (T) (1.0 / std::sqrt((double) n)));
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

template <typename T> int fft3d_generic(T *field, int n, int dir, cudaStream_t st) {
  Plan plan;
  if ((n & 1) || !make_plan(n, plan)) return fail("the mixed-radix transform needs an even grid size");
  const C2<T> *tw = nullptr;
  if (int rc = twiddles<T>(n, &tw)) return rc;
  if (dir < 0) {
    if (int rc = run_pass<T, AX_Z_R2C>(field, plan, tw, dir, st, "fft_generic_z")) return rc;
    if (int rc = run_pass<T, AX_Y>(field, plan, tw, dir, st, "fft_generic_y")) return rc;
    return run_pass<T, AX_X>(field, plan, tw, dir, st, "fft_generic_x");
  }
  if (int rc = run_pass<T, AX_X>(field, plan, tw, dir, st, "fft_generic_x")) return rc;
  if (int rc = run_pass<T, AX_Y>(field, plan, tw, dir, st, "fft_generic_y")) return rc;
  return run_pass<T, AX_Z_C2R>(field, plan, tw, dir, st, "fft_generic_z");
}
template int fft3d_generic<float>(float *, int, int, cudaStream_t);
template int fft3d_generic<double>(double *, int, int, cudaStream_t);

}  // namespace fftgen

namespace fft {
template <typename T> int fft3d(T *field, int n, int dir, cudaStream_t st);
// any supported size: the tuned passes for powers of two, the mixed-radix passes otherwise
template <typename T> int fft3d_any(T *field, int n, int dir, cudaStream_t st) {
  return is_pow2(n) ? fft3d<T>(field, n, dir, st) : fftgen::fft3d_generic<T>(field, n, dir, st);
}
template int fft3d_any<float>(float *, int, int, cudaStream_t);
template int fft3d_any<double>(double *, int, int, cudaStream_t);
}  // namespace fft
}  // namespace gx

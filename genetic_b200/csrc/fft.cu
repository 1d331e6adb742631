// 3-D real<->half-complex FFT on the reference's padded layout, as three hand-written 1-D passes.
//
// Replaces tools::numerics::fourier::FieldFourierManager<T,T>::performTransform
// (/root/reference/genetIC/src/tools/numerics/fourier.hpp:479-518) including the pad/unpad sweeps
// (:368-405), the divide-by-N^1.5 sweep (:490,512-513) and, for the inverse, the Hermitian mirror
// (:331-357, see spectral.cu). Layout in HBM: N x N rows of (N/2+1) complex == N+2 reals, z fastest
// (grid.hpp:610-626); real-space data keeps the padded pitch on the device (the reference's unpadded
// N^3 layout is produced/consumed only by the host upload/download copies, capi.cu).
//
//   forward : z-pass r2c (rows, contiguous) -> y-pass (stride nc) -> x-pass (stride n*nc), scale on last
//   inverse : x-pass -> y-pass -> z-pass c2r (rows), scale on last
//
// Each pass reads and writes the field exactly once (algorithmic traffic 2*B per pass, 6*B per 3-D
// transform, B = bytes of the padded field).
#include "fft.cuh"

namespace gx {
namespace fft {

// ------------------------------------------------------------------------------------------------
// twiddle tables: for every power of two n in [2, 4096], n entries exp(-2 pi i j / n) at offset n.
// ------------------------------------------------------------------------------------------------
static double2 *g_tw_d = nullptr;
static float2 *g_tw_f = nullptr;
static int g_tw_device = -1;
static const int kTwTotal = 8192;

static void exact_cossin(long long j, long long n, long double &c, long double &s) {
  // reduce to the first octant so table entries are symmetric to the last bit
  j %= n;
  long long q = (8 * j) / n;  // octant
  long double a;
  const long double tau = 6.283185307179586476925286766559005768394L;
  switch (q) {
    case 0: a = tau * j / n; c = cosl(a); s = sinl(a); break;
    case 1: a = tau * (n / 4.0L - j) / n; c = sinl(a); s = cosl(a); break;
    case 2: a = tau * (j - n / 4.0L) / n; c = -sinl(a); s = cosl(a); break;
    case 3: a = tau * (n / 2.0L - j) / n; c = -cosl(a); s = sinl(a); break;
    case 4: a = tau * (j - n / 2.0L) / n; c = -cosl(a); s = -sinl(a); break;
    case 5: a = tau * (3 * n / 4.0L - j) / n; c = -sinl(a); s = -cosl(a); break;
    case 6: a = tau * (j - 3 * n / 4.0L) / n; c = sinl(a); s = -cosl(a); break;
    default: a = tau * (n - j) / n; c = cosl(a); s = -sinl(a); break;
  }
}

int ensure_twiddles() {
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_tw_d != nullptr && dev == g_tw_device) return 0;
  std::vector<double2> hd(kTwTotal);
  std::vector<float2> hf(kTwTotal);
  for (int i = 0; i < kTwTotal; ++i) { hd[i] = make_double2(1, 0); hf[i] = make_float2(1, 0); }
  for (int n = 2; n <= 4096; n *= 2) {
    for (int j = 0; j < n; ++j) {
      long double c, s;
      exact_cossin(j, n, c, s);
      hd[n + j] = make_double2((double) c, (double) -s);
      hf[n + j] = make_float2((float) c, (float) -s);
    }
  }
  GX_CUDA_CHECK(cudaMalloc(&g_tw_d, kTwTotal * sizeof(double2)));
  GX_CUDA_CHECK(cudaMalloc(&g_tw_f, kTwTotal * sizeof(float2)));
  GX_CUDA_CHECK(cudaMemcpy(g_tw_d, hd.data(), kTwTotal * sizeof(double2), cudaMemcpyHostToDevice));
  GX_CUDA_CHECK(cudaMemcpy(g_tw_f, hf.data(), kTwTotal * sizeof(float2), cudaMemcpyHostToDevice));
  g_tw_device = dev;
  return 0;
}

template <> const double2 *twiddle_table<double>(int n) { return g_tw_d + n; }
template <> const float2 *twiddle_table<float>(int n) { return g_tw_f + n; }

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
template <typename T, int N> static int fft3d_n(T *data, int dir, cudaStream_t st) {
  using C = cplx<T>;
  const T scale = (T) (1.0 / pow((double) N, 1.5));
  C *cdata = reinterpret_cast<C *>(data);
  const int nc = N / 2 + 1;
  if (dir < 0) {
    int rc = launch_rows_r2c<T, N>(data, (long) N * N, st);
    if (rc) return rc;
    rc = launch_cols<T, N, -1, PassY>(cdata, cdata, N, N, nc, OpIdentity<T>(), OpIdentity<T>(), st);
    if (rc) return rc;
    return launch_cols<T, N, -1, PassX>(cdata, cdata, N, N, nc, OpIdentity<T>(), OpScale<T>(scale), st);
  } else {
    int rc = launch_cols<T, N, 1, PassX>(cdata, cdata, N, N, nc, OpIdentity<T>(), OpIdentity<T>(), st);
    if (rc) return rc;
    rc = launch_cols<T, N, 1, PassY>(cdata, cdata, N, N, nc, OpIdentity<T>(), OpIdentity<T>(), st);
    if (rc) return rc;
    StoreRealScaled<T> store;
    store.data = data; store.scale = scale; store.n = N;
    return launch_rows_c2r<T, N>(data, 0, (long) N * N, store, st);
  }
}

template <typename T> int fft3d(T *data, int n, int dir, cudaStream_t st) {
  int rc = ensure_twiddles();
  if (rc) return rc;
  switch (n) {
#define GX_CASE(N_) case N_: return fft3d_n<T, N_>(data, dir, st);
    GX_CASE(4) GX_CASE(8) GX_CASE(16) GX_CASE(32) GX_CASE(64) GX_CASE(128) GX_CASE(256) GX_CASE(512)
    GX_CASE(1024) GX_CASE(2048) GX_CASE(4096)
#undef GX_CASE
    default:
      return fail("gx_fft: grid size " + std::to_string(n) + " unsupported (power of two in [4,4096] required)");
  }
}

template int fft3d<float>(float *, int, int, cudaStream_t);
template int fft3d<double>(double *, int, int, cudaStream_t);

}  // namespace fft
}  // namespace gx

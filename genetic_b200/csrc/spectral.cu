// Spectral element-wise kernels + their C-ABI entry points.
#include "../../include/genetic_b200.h"
#include "spectral.cuh"
#include "spectrum_host.h"
#include <vector>
#include <cstdlib>

namespace gx {

struct NoInit { __device__ __forceinline__ void init() const {} };

// ---- ensureFourierModesAreMirrored (fourier.hpp:331-357) ---------------------------------------------
// One thread per (plane in {kz=0, kz=n/2}, kx in [0, n/2], ky in [0, n)).
template <typename T> __global__ void mirror_kernel(cplx<T> *f, int n, int nc) {
  const int h = n >> 1;
  const int total = 2 * (h + 1) * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ky = i % n;
    const int kx = (i / n) % (h + 1);
    const int iz = (i / (n * (h + 1))) ? h : 0;
    const size_t src = ((size_t) kx * n + ky) * nc + iz;
    if (kx == 0 || kx == h) {
      // within-row: +ky -> -ky for 0 < ky < h; the self-conjugate entries get their imaginary part negated,
      // exactly as the reference's sequential descending-ky loop leaves them
      if (ky == 0 || ky == h) { f[src].y = -f[src].y; }
      else if (ky < h) { cplx<T> v = f[src]; v.y = -v.y; f[((size_t) kx * n + (n - ky)) * nc + iz] = v; }
    } else {
      cplx<T> v = f[src];
      v.y = -v.y;
      f[((size_t) (n - kx) * n + ((n - ky) % n)) * nc + iz] = v;
    }
  }
}

template <typename T> int mirror(T *field, int n, cudaStream_t st) {
  const int nc = n / 2 + 1;
  const int total = 2 * (n / 2 + 1) * n;
  GX_TIMED("mirror", st);
  mirror_kernel<T><<<(total + 255) / 256, 256, 0, st>>>(reinterpret_cast<cplx<T> *>(field), n, nc);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}
template int mirror<float>(float *, int, cudaStream_t);
template int mirror<double>(double *, int, cudaStream_t);

// ---- spectrum table staging ------------------------------------------------------------------------
struct SpecInit {
  SpecDev sp;
  __device__ __forceinline__ void init() const {
    extern __shared__ double s_tab[];
    for (int i = threadIdx.x; i < 3 * sp.m; i += blockDim.x) s_tab[i] = sp.tab[i];
    __syncthreads();
  }
  __device__ __forceinline__ const double *lx() const { extern __shared__ double s_tab[]; return s_tab; }
};

// ---- getPowerSpectrumForGridUncached (camb.hpp:101-119) ------------------------------------------------
template <typename T> struct CovarianceF : SpecInit {
  cplx<T> *out; int n; T kmin, norm;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    cplx<T> v = mk<T>(0, 0);
    if (slot_is_independent(n, ix, iy, iz)) {
      T kx, ky, kz, k2;
      slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
      const double *t = lx();
      v.x = mul_rn(spectrum_power<T>(sp, t, t + sp.m, t + 2 * sp.m, sqrt_rn(k2)), norm);
    }
    out[s] = v;
  }
};

// ---- Field::applyTransferFunction with the covariance evaluated on the fly (field.hpp:394-412) ------------
// divide_k2: additionally leave delta_k / k^2 (0 at k = 0) -- the quotient every Zeldovich component starts from
// (zeldovich.hpp:170-176 divides per component; the rounded quotient is the same number, so it is formed once here
// instead of three times inside the FFT passes).
template <typename T> struct ApplyTransferF : SpecInit {
  cplx<T> *f; int n; T kmin, norm; double power; int divide_k2;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    T kx, ky, kz, k2;
    slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
    const double *t = lx();
    T spec = mul_rn(spectrum_power<T>(sp, t, t + sp.m, t + 2 * sp.m, sqrt_rn(k2)), norm);
    spec = spec_pow<T>(spec, power);
    cplx<T> v = f[s];
    v = mk<T>(mul_rn(v.x, spec), mul_rn(v.y, spec));
    if (divide_k2) v = (k2 != (T) 0) ? mk<T>(div_rn(v.x, k2), div_rn(v.y, k2)) : mk<T>(0, 0);
    f[s] = v;
  }
};

// ---- the same through a table over the integer |k|^2 ---------------------------------------------------------------
// C(k) only depends on m = ix^2 + iy^2 + iz^2 (signed mode numbers), at most 3 (n/2)^2 + 1 distinct values (25 MB of
// doubles at n = 2048: resident in the 126 MB L2). Evaluating log / spline / exp / pow once per m instead of once per
// mode turns the sweep from FP64-bound (53 % FP64 pipe, 1.4 TB/s at 2048^3, profiles/r1s3_*) into a streaming pass.
// The table is rebuilt by every call (a few hundred microseconds): nothing is cached between calls.
// Rounding: the reference forms k^2 from the three rounded products (slot_k); the table uses k^2 = m * kmin^2. The two
// differ by a few ulp, i.e. ~1e-16 relative in the factor -- the north star's tolerance is 1e-10 (the in-kernel
// evaluation above already differs from glibc's pow/log/exp by an ulp).
template <typename T> struct TransferTableF : SpecInit {
  T *tab; long mcount; T kmin, norm; double power; int divide_k2;
  __device__ __forceinline__ void build(long m) const {
    const T k2 = mul_rn((T) m, mul_rn(kmin, kmin));
    const double *t = lx();
    T spec = mul_rn(spectrum_power<T>(sp, t, t + sp.m, t + 2 * sp.m, sqrt_rn(k2)), norm);
    spec = spec_pow<T>(spec, power);
    if (divide_k2) spec = (m != 0) ? div_rn(spec, k2) : (T) 0;
    tab[m] = spec;
  }
};
template <typename T> __global__ void __launch_bounds__(256) transfer_table_kernel(TransferTableF<T> f) {
  f.init();
  for (long m = (long) blockIdx.x * blockDim.x + threadIdx.x; m < f.mcount; m += (long) gridDim.x * blockDim.x) f.build(m);
}
template <typename T> struct ApplyTableF : NoInit {
  cplx<T> *f; const T *__restrict__ tab; int n;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    const int a = signed_mode(ix, n), b = signed_mode(iy, n);
    const T w = __ldg(tab + ((long) a * a + (long) b * b + (long) iz * iz));
    cplx<T> v = f[s];
    f[s] = mk<T>(mul_rn(v.x, w), mul_rn(v.y, w));
  }
};

template <typename T> struct ApplyTransferFieldF : NoInit {
  cplx<T> *f; const cplx<T> *cov; int n; double power;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    T spec = spec_pow<T>(cov[s].x, power);
    cplx<T> v = f[s];
    f[s] = mk<T>(mul_rn(v.x, spec), mul_rn(v.y, spec));
  }
};

// ---- filters (filter.hpp) ---------------------------------------------------------------------------
struct FilterDesc { int nf; gx_filter_factor f[16]; };

template <typename T> __device__ __forceinline__ T filter_eval(const FilterDesc &d, T k) {
  T r = (T) 1;
  for (int i = 0; i < d.nf; ++i) {
    const gx_filter_factor &q = d.f[i];
    T v;
    if (q.kind == GX_FILTER_CONST) v = (T) q.kcut;
    else {
      // LowPassFermiFilter::operator() : 1./(1.+exp((k-kcut)/temperature)) in T
      const T kc = (T) q.kcut, tp = (T) q.temperature;
      const T lp = (T) 1. / ((T) 1. + (T) exp((double) ((k - kc) / tp)));
      if (q.kind == GX_FILTER_FERMI_LP) v = lp;
      else if (q.kind == GX_FILTER_FERMI_HP_COV) v = sqrt_rn((T) 1. - lp * lp);
      else v = (T) 1. - lp;
    }
    for (int p = 0; p < q.power; ++p) r *= v;
  }
  return r;
}

template <typename T> struct ApplyFilterF : NoInit {
  cplx<T> *f; int n; T kmin; FilterDesc d;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    T kx, ky, kz, k2;
    slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
    // field.hpp:743: k = sqrt(double(kx*kx+ky*ky+kz*kz)) -> CoordinateType
    const T k = (T) sqrt((double) k2);
    const T w = filter_eval<T>(d, k);
    cplx<T> v = f[s];
    f[s] = mk<T>(mul_rn(v.x, w), mul_rn(v.y, w));
  }
};

// ---- ICGenerator::reverseSmallK (ic.hpp:1690-1708): delta_k <- -delta_k where 0 < k^2 < kmax^2 ---------------------------
template <typename T> struct ReverseSmallKF : NoInit {
  cplx<T> *f; int n; T kmin, k2max;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    T kx, ky, kz, k2;
    slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
    if (k2 < k2max && k2 != (T) 0) {
      const cplx<T> v = f[s];
      f[s] = mk<T>(-v.x, -v.y);
    }
  }
};

// ---- enforceUnitVariance (multilevelfield.hpp:498-510) ---------------------------------------------------
template <typename T> struct UnitVarianceF : NoInit {
  cplx<T> *f; int n;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    cplx<T> v = f[s];
    const T a = (T) hypot((double) v.x, (double) v.y);
    if (a == (T) 0) f[s] = mk<T>(0, 0);
    else f[s] = mk<T>(div_rn(v.x, a), div_rn(v.y, a));
  }
};

// ---- Zeldovich operator, unfused form (zeldovich.hpp:156-181 through fourier.hpp:253-274) ---------------
template <typename T> struct ZeldovichF : NoInit {
  const cplx<T> *delta; cplx<T> *psi; int n, axis; T kmin;
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    cplx<T> r = mk<T>(0, 0);
    if (slot_is_independent(n, ix, iy, iz)) {
      T kx, ky, kz, k2;
      slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
      const int h = n >> 1;
      const int ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
      const T ka = axis == 0 ? kx : (axis == 1 ? ky : kz);
      if (ia != h && k2 != (T) 0) {
        const cplx<T> d = delta[s];
        r.x = mul_rn(div_rn(-d.y, k2), ka);
        r.y = mul_rn(div_rn(d.x, k2), ka);
      }
    }
    psi[s] = r;
  }
};

template <typename T> static T kmin_of(double grid_length) { return (T) (2. * M_PI / (double) (T) grid_length); }

}  // namespace gx

using namespace gx;

#define GX_DISPATCH(dtype, CALL)                                      \
  do {                                                                \
    if ((dtype) == GX_F64) { using T = double; return CALL; }         \
    if ((dtype) == GX_F32) { using T = float; return CALL; }          \
    return gx::fail("bad dtype");                                     \
  } while (0)

static int check_n(int n) {
  if (n < 4 || n > 4096 || (n & 1)) return gx::fail("grid size must be even and in [4,4096]");  // sweeps: any even size
  return 0;
}

template <typename T> static SpecInit make_spec_init(const gx_spectrum *s) {
  SpecInit si;
  si.sp = gx_spectrum_device_view(s);
  return si;
}

template <typename T> static int covariance_field_t(void *out, int n, double L, const gx_spectrum *s, cudaStream_t st) {
  CovarianceF<T> f;
  f.sp = gx_spectrum_device_view(s);
  f.out = (cplx<T> *) out; f.n = n; f.kmin = kmin_of<T>(L); f.norm = (T) gx_spectrum_grid_norm(s, L, n);
  return launch_slot_sweep(n, f, st, 3 * f.sp.m * sizeof(double), "covariance_field");
}

static void *g_transfer_tab = nullptr;   // grow-only workspace for the |k|^2 table (values are rebuilt by every call)
static size_t g_transfer_tab_bytes = 0;

static int ensure_transfer_tab(size_t bytes) {
  if (g_transfer_tab_bytes < bytes) {
    if (g_transfer_tab) { GX_CUDA_CHECK(cudaDeviceSynchronize()); GX_CUDA_CHECK(cudaFree(g_transfer_tab)); g_transfer_tab = nullptr; }
    GX_CUDA_CHECK(cudaMalloc(&g_transfer_tab, bytes));
    g_transfer_tab_bytes = bytes;
  }
  return 0;
}

// table[m] = C(k)^power [/ k^2] for m = ix^2 + iy^2 + iz^2; rebuilt by every call, only the allocation is kept
template <typename T>
static int build_transfer_table_t(int n, double L, const gx_spectrum *s, double power, int divide_k2, cudaStream_t st) {
  const long h = n / 2, mcount = 3 * h * h + 1;
  if (int rc = ensure_transfer_tab(mcount * sizeof(T))) return rc;
  TransferTableF<T> tf;
  tf.sp = gx_spectrum_device_view(s);
  tf.tab = (T *) g_transfer_tab; tf.mcount = mcount; tf.kmin = kmin_of<T>(L); tf.norm = (T) gx_spectrum_grid_norm(s, L, n);
  tf.power = power; tf.divide_k2 = divide_k2;
  {
    GX_TIMED("transfer_table", st);
    long blocks = (mcount + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const size_t smem = 3 * tf.sp.m * sizeof(double);
    if (smem > 48 * 1024) {
      if (smem > 200 * 1024) return fail("transfer-function table too long for the in-kernel spline (more than ~8500 rows)");
      GX_CUDA_CHECK(cudaFuncSetAttribute(transfer_table_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    transfer_table_kernel<T><<<(unsigned) blocks, 256, smem, st>>>(tf);
    count_launch();
  }
  GX_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int apply_transfer_t(void *field, int n, double L, const gx_spectrum *s, double power, cudaStream_t st,
                            int divide_k2 = 0, int iy0 = 0, int ny = -1) {
  static int use_table = -1;
  if (use_table < 0) { const char *e = getenv("GX_TRANSFER_TABLE"); use_table = (e && e[0] == '0') ? 0 : 1; }
  if (use_table) {
    if (int rc = build_transfer_table_t<T>(n, L, s, power, divide_k2, st)) return rc;
    ApplyTableF<T> f;
    f.f = (cplx<T> *) field; f.tab = (const T *) g_transfer_tab; f.n = n;
    return launch_slot_sweep(n, f, st, 0, "apply_transfer", iy0, ny);
  }
  ApplyTransferF<T> f;
  f.sp = gx_spectrum_device_view(s);
  f.f = (cplx<T> *) field; f.n = n; f.kmin = kmin_of<T>(L); f.norm = (T) gx_spectrum_grid_norm(s, L, n); f.power = power;
  f.divide_k2 = divide_k2;
  return launch_slot_sweep(n, f, st, 3 * f.sp.m * sizeof(double), "apply_transfer", iy0, ny);
}

template <typename T> static int apply_transfer_field_t(void *field, const void *cov, int n, double power, cudaStream_t st) {
  ApplyTransferFieldF<T> f;
  f.f = (cplx<T> *) field; f.cov = (const cplx<T> *) cov; f.n = n; f.power = power;
  return launch_slot_sweep(n, f, st);
}

// table[m] = f(|k|) for m = ix^2 + iy^2 + iz^2 (|k| = sqrt(m kMin^2)): the exponentials of the Fermi factors are evaluated
// once per distinct |k|^2 (3 (n/2)^2 + 1 values) instead of once per mode, and the sweep becomes a pure HBM pass
template <typename T> struct FilterTableF {
  T *tab; long mcount; T kmin; FilterDesc d;
};
template <typename T> __global__ void __launch_bounds__(256) filter_table_kernel(FilterTableF<T> f) {
  for (long m = (long) blockIdx.x * blockDim.x + threadIdx.x; m < f.mcount; m += (long) gridDim.x * blockDim.x) {
    const T k2 = mul_rn((T) m, mul_rn(f.kmin, f.kmin));
    f.tab[m] = filter_eval<T>(f.d, (T) sqrt((double) k2));
  }
}

template <typename T>
static int apply_filter_t(void *field, int n, double L, const gx_filter_factor *fac, int nf, cudaStream_t st) {
  if (nf < 0 || nf > 16) return gx::fail("gx_apply_filter: at most 16 factors");
  static int use_table = -1;
  if (use_table < 0) { const char *e = getenv("GX_FILTER_TABLE"); use_table = (e && e[0] == '0') ? 0 : 1; }
  if (use_table) {
    const long h = n / 2, mcount = 3 * h * h + 1;
    if (int rc = ensure_transfer_tab(mcount * sizeof(T))) return rc;
    FilterTableF<T> tf;
    tf.tab = (T *) g_transfer_tab; tf.mcount = mcount; tf.kmin = kmin_of<T>(L); tf.d.nf = nf;
    for (int i = 0; i < nf; ++i) tf.d.f[i] = fac[i];
    {
      GX_TIMED("filter_table", st);
      long blocks = (mcount + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      filter_table_kernel<T><<<(unsigned) blocks, 256, 0, st>>>(tf);
      count_launch();
    }
    GX_LAUNCH_CHECK();
    ApplyTableF<T> f;
    f.f = (cplx<T> *) field; f.tab = (const T *) g_transfer_tab; f.n = n;
    return launch_slot_sweep(n, f, st, 0, "apply_filter");
  }
  ApplyFilterF<T> f;
  f.f = (cplx<T> *) field; f.n = n; f.kmin = kmin_of<T>(L); f.d.nf = nf;
  for (int i = 0; i < nf; ++i) f.d.f[i] = fac[i];
  return launch_slot_sweep(n, f, st);
}

template <typename T> static int reverse_small_k_t(void *field, int n, double L, double kmax, cudaStream_t st) {
  ReverseSmallKF<T> f;
  f.f = (cplx<T> *) field; f.n = n; f.kmin = kmin_of<T>(L); f.k2max = (T) kmax * (T) kmax;
  return launch_slot_sweep(n, f, st, 0, "reverse_small_k");
}

template <typename T> static int unit_variance_t(void *field, int n, cudaStream_t st) {
  UnitVarianceF<T> f;
  f.f = (cplx<T> *) field; f.n = n;
  return launch_slot_sweep(n, f, st);
}

template <typename T>
static int zeldovich_fourier_t(const void *delta, void *psi, int n, double L, int axis, cudaStream_t st) {
  ZeldovichF<T> f;
  f.delta = (const cplx<T> *) delta; f.psi = (cplx<T> *) psi; f.n = n; f.axis = axis; f.kmin = kmin_of<T>(L);
  return launch_slot_sweep(n, f, st, 0, "zeldovich_op");
}

namespace gx {
// ---- cosmology::dumpPowerSpectrum (cosmology/parameters.hpp:112-205): the binning reduction on the device ------------------
// The reference visits every (ix, iy, iz) in [-n/2, n/2]^3 through getFourierCoefficient -- each stored slot (a, b, c >= 0)
// once for +iz and once (as the conjugate, same modulus) for -iz, and twice more per index that sits on the Nyquist plane
// (ix = -n/2 and +n/2 are the same slot) -- and adds |delta_k|^2, Re C(k), k and 1 to the logarithmic bin of
// k = sqrt(ix^2 + iy^2 + iz^2) * 2 pi / L if kmin <= k < kmax. Here: one thread per stored slot with that multiplicity as its
// weight, bins in shared memory per CTA, double atomics into the result. bins[4][nbins] = count, sum k, sum |delta|^2, sum C.
template <typename T>
__global__ void ps_bins_kernel(const cplx<T> *f, const cplx<T> *cov, int n, T kw, const int *__restrict__ bin_of_m, int nbins,
                               double *bins) {
  extern __shared__ double sb[];  // [4][nbins]
  for (int i = threadIdx.x; i < 4 * nbins; i += blockDim.x) sb[i] = 0.0;
  __syncthreads();
  const int h = n >> 1, nc = h + 1;
  const long rows = (long) n * n;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int a = (int) (row / n), b = (int) (row % n);
    const int sa = signed_mode(a, n), sb2 = signed_mode(b, n);
    const int wab = (a == h ? 2 : 1) * (b == h ? 2 : 1);
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
      const int m = sa * sa + sb2 * sb2 + c * c;
      const int idx = bin_of_m[m];  // host-made with the C library's log10: a bin edge is one ulp of log10 wide (below)
      if (idx < 0) continue;
      const T kfft = (T) sqrt((double) m);
      const T k = kfft * kw;
      const double w = (double) (wab * (c == 0 ? 1 : 2));
      const cplx<T> v = f[row * nc + c];
      T vabs = (T) hypot((double) v.x, (double) v.y);  // std::abs(complex)
      vabs *= vabs;
      atomicAdd(&sb[idx], w);
      atomicAdd(&sb[nbins + idx], w * (double) k);
      atomicAdd(&sb[2 * nbins + idx], w * (double) vabs);
      if (cov) atomicAdd(&sb[3 * nbins + idx], w * (double) cov[row * nc + c].x);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * nbins; i += blockDim.x)
    if (sb[i] != 0.0) atomicAdd(&bins[i], sb[i]);
}

template <typename T>
static int power_spectrum_bins_t(void *field, void *cov, int n, double L, int nbins, double *host_bins, cudaStream_t st) {
  // field.ensureFourierModesAreMirrored(); P0.ensureFourierModesAreMirrored();  (:121-122)
  if (int rc = mirror<T>((T *) field, n, st)) return rc;
  if (cov) if (int rc = mirror<T>((T *) cov, n, st)) return rc;
  const T boxLength = (T) L;
  const T kmax = (T) M_PI / boxLength * (T) n, kmin = 2.0f * (T) M_PI / boxLength;
  const T dklog = (T) log10((double) (kmax / kmin)) / (T) nbins, kw = 2.0f * (T) M_PI / boxLength;
  // The bin of a mode depends on m = ix^2 + iy^2 + iz^2 only. Modes sit exactly on bin edges (k / kmin = 2 when (n/2)^(j/100) = 2),
  // where the truncation of log10(k / kmin) / dklog turns on the last bit of log10: the table is made on the host with the C
  // library's log10, the one the reference calls, instead of the device's (1 ulp) -- found by the reference's test_01d on the
  // B200, whose power-spectrum file moved 12 modes between two bins.
  const long h = n / 2, mcount = 3 * h * h + 1;
  std::vector<int> bin_of_m((size_t) mcount, -1);
  const T inv_dklog = (T) (1.0f / dklog);
  for (long m = 0; m < mcount; ++m) {
    const T kfft = (T) std::sqrt((double) m);
    const T k = kfft * kw;
    if (!(k >= kmin && k < kmax)) continue;
    const int idx = (int) (inv_dklog * std::log10(k / kmin));
    if (idx >= 0 && idx < nbins) bin_of_m[(size_t) m] = idx;
  }
  int *d_bin_of_m = nullptr;
  GX_CUDA_CHECK(cudaMalloc(&d_bin_of_m, sizeof(int) * mcount));
  GX_CUDA_CHECK(cudaMemcpyAsync(d_bin_of_m, bin_of_m.data(), sizeof(int) * mcount, cudaMemcpyHostToDevice, st));
  double *d_bins = nullptr;
  GX_CUDA_CHECK(cudaMalloc(&d_bins, sizeof(double) * 4 * nbins));
  GX_CUDA_CHECK(cudaMemsetAsync(d_bins, 0, sizeof(double) * 4 * nbins, st));
  const long rows = (long) n * n;
  const int grid = (int) (rows < 148 * 8 ? rows : 148 * 8);
  ps_bins_kernel<T><<<grid, 256, sizeof(double) * 4 * nbins, st>>>((const cplx<T> *) field, (const cplx<T> *) cov, n, kw, d_bin_of_m,
                                                                 nbins, d_bins);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_bins, d_bins, sizeof(double) * 4 * nbins, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_bins);
  cudaFree(d_bin_of_m);
  if (e != cudaSuccess) return fail(std::string("gx_power_spectrum_bins: ") + cudaGetErrorString(e));
  return 0;
}

void release_spectral_workspace() {
  if (g_transfer_tab) { cudaDeviceSynchronize(); cudaFree(g_transfer_tab); }
  g_transfer_tab = nullptr; g_transfer_tab_bytes = 0;
}
// the |k|^2 table alone, for consumers that multiply while they write (the noise delivery in gx_unigrid_ic)
int build_transfer_table(int n, int dtype, double L, const gx_spectrum *s, double power, int divide_k2, cudaStream_t st,
                         const void **table) {
  if (check_n(n)) return 1;
  int rc;
  if (dtype == GX_F64) rc = build_transfer_table_t<double>(n, L, s, power, divide_k2, st);
  else if (dtype == GX_F32) rc = build_transfer_table_t<float>(n, L, s, power, divide_k2, st);
  else return fail("bad dtype");
  *table = g_transfer_tab;
  return rc;
}
int apply_transfer_ex(void *field, int n, int dtype, double L, const gx_spectrum *s, double power, int divide_k2,
                      cudaStream_t st, int iy0, int ny) {
  if (check_n(n)) return 1;
  if (dtype == GX_F64) return apply_transfer_t<double>(field, n, L, s, power, st, divide_k2, iy0, ny);
  if (dtype == GX_F32) return apply_transfer_t<float>(field, n, L, s, power, st, divide_k2, iy0, ny);
  return fail("bad dtype");
}
template <typename T> static int unit_variance_slab_t(void *field, int n, int iy0, int ny, cudaStream_t st) {
  UnitVarianceF<T> f;
  f.f = (cplx<T> *) field; f.n = n;
  return launch_slot_sweep(n, f, st, 0, "unit_variance", iy0, ny);
}
int unit_variance_slab(void *field, int n, int dtype, int iy0, int ny, cudaStream_t st) {
  if (dtype == GX_F64) return unit_variance_slab_t<double>(field, n, iy0, ny, st);
  if (dtype == GX_F32) return unit_variance_slab_t<float>(field, n, iy0, ny, st);
  return fail("bad dtype");
}
}  // namespace gx

extern "C" {

int gx_mirror(void *field, int n, int dtype, void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, gx::mirror<T>((T *) field, n, (cudaStream_t) stream));
}

int gx_power_spectrum_bins(void *field, void *cov_field, int n, int dtype, double grid_length, int nbins, double *host_bins,
                           void *stream) {
  if (check_n(n)) return 1;
  if (nbins < 1 || nbins > 256) return fail("gx_power_spectrum_bins: 1 <= nbins <= 256");
  if (!field || !host_bins) return fail("gx_power_spectrum_bins: null argument");
  GX_DISPATCH(dtype, power_spectrum_bins_t<T>(field, cov_field, n, grid_length, nbins, host_bins, (cudaStream_t) stream));
}

int gx_covariance_field(void *out_field, int n, int dtype, double grid_length, const gx_spectrum *s, void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, covariance_field_t<T>(out_field, n, grid_length, s, (cudaStream_t) stream));
}

int gx_apply_transfer(void *field, int n, int dtype, double grid_length, const gx_spectrum *s, double power,
                      void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, apply_transfer_t<T>(field, n, grid_length, s, power, (cudaStream_t) stream));
}

int gx_apply_transfer_field(void *field, const void *cov_field, int n, int dtype, double power, void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, apply_transfer_field_t<T>(field, cov_field, n, power, (cudaStream_t) stream));
}

int gx_apply_filter(void *field, int n, int dtype, double grid_length, const gx_filter_factor *factors, int nfactors,
                    void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, apply_filter_t<T>(field, n, grid_length, factors, nfactors, (cudaStream_t) stream));
}

int gx_reverse_small_k(void *field, int n, int dtype, double grid_length, double kmax, void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, reverse_small_k_t<T>(field, n, grid_length, kmax, (cudaStream_t) stream));
}

int gx_unit_variance(void *field, int n, int dtype, void *stream) {
  if (check_n(n)) return 1;
  GX_DISPATCH(dtype, unit_variance_t<T>(field, n, (cudaStream_t) stream));
}

int gx_zeldovich_fourier(const void *delta_k, void *psi_k, int n, int dtype, double grid_length, int axis,
                         void *stream) {
  if (check_n(n)) return 1;
  if (axis < 0 || axis > 2) return gx::fail("axis must be 0, 1 or 2");
  GX_DISPATCH(dtype, zeldovich_fourier_t<T>(delta_k, psi_k, n, grid_length, axis, (cudaStream_t) stream));
}

}  // extern "C"

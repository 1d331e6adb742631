// TEST INFRASTRUCTURE: the FFTW3 entry points of oracle/shims/fftw3.h as a plain radix-2 transform (power-of-two sizes),
// OpenMP over lines. In-place 3-D transforms on FFTW's layout: real data [n0][n1][2(n2/2+1)] (padded), complex data
// [n0][n1][n2/2+1]; forward = exp(-2 pi i ..), unnormalised; c2r ignores the imaginary parts of the self-conjugate bins.
#include "fftw3.h"
#include <cmath>
#include <complex>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace {
template <typename T> struct Tw {
  int n = 0;
  std::vector<std::complex<T>> w;  // exp(-2 pi i j / n), j < n/2
  std::vector<int> rev;
  void init(int n_) {
    if (n == n_) return;
    n = n_;
    if (n & (n - 1)) throw std::runtime_error("fft shim: size must be a power of two");
    w.resize(n / 2 > 0 ? n / 2 : 1);
    for (int j = 0; j < n / 2; ++j) {
      const long double a = -2.0L * M_PIl * j / n;
      w[j] = std::complex<T>((T) cosl(a), (T) sinl(a));
    }
    rev.assign(n, 0);
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
      int r = 0;
      for (int b = 0; b < bits; ++b) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
      rev[i] = r;
    }
  }
};

// in-place complex FFT of length t.n on a contiguous line; sign -1 forward, +1 backward
template <typename T> void fft_line(std::complex<T> *a, const Tw<T> &t, int sign) {
  const int n = t.n;
  for (int i = 0; i < n; ++i) if (i < t.rev[i]) std::swap(a[i], a[t.rev[i]]);
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len)
      for (int j = 0; j < half; ++j) {
        std::complex<T> w = t.w[j * step];
        if (sign > 0) w = std::conj(w);
        const std::complex<T> u = a[i + j], v = a[i + j + half] * w;
        a[i + j] = u + v;
        a[i + j + half] = u - v;
      }
  }
}

template <typename T> struct Plan {
  int n0, n1, n2, kind;  // kind 0: r2c, 1: c2r, 2: complex forward, 3: complex backward
  T *data;
};

template <typename T> void execute(const Plan<T> &p) {
  using C = std::complex<T>;
  const int n0 = p.n0, n1 = p.n1, n2 = p.n2;
  if (p.kind >= 2) {  // complex 3-D, [n0][n1][n2]
    C *d = reinterpret_cast<C *>(p.data);
    const int sign = p.kind == 2 ? -1 : 1;
    Tw<T> t0, t1, t2;
    t0.init(n0); t1.init(n1); t2.init(n2);
#pragma omp parallel
    {
      std::vector<C> line(std::max(n0, std::max(n1, n2)));
#pragma omp for collapse(2)
      for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) fft_line(d + ((size_t) i * n1 + j) * n2, t2, sign);
#pragma omp for collapse(2)
      for (int i = 0; i < n0; ++i) for (int k = 0; k < n2; ++k) {
        for (int j = 0; j < n1; ++j) line[j] = d[((size_t) i * n1 + j) * n2 + k];
        fft_line(line.data(), t1, sign);
        for (int j = 0; j < n1; ++j) d[((size_t) i * n1 + j) * n2 + k] = line[j];
      }
#pragma omp for collapse(2)
      for (int j = 0; j < n1; ++j) for (int k = 0; k < n2; ++k) {
        for (int i = 0; i < n0; ++i) line[i] = d[((size_t) i * n1 + j) * n2 + k];
        fft_line(line.data(), t0, sign);
        for (int i = 0; i < n0; ++i) d[((size_t) i * n1 + j) * n2 + k] = line[i];
      }
    }
    return;
  }
  const int nc = n2 / 2 + 1;
  C *c = reinterpret_cast<C *>(p.data);
  Tw<T> t0, t1, t2;
  t0.init(n0); t1.init(n1); t2.init(n2);
  const int sign = p.kind == 0 ? -1 : 1;
#pragma omp parallel
  {
    std::vector<C> line(std::max(n0, std::max(n1, n2)));
    if (p.kind == 0) {  // rows first: real line -> n2/2+1 complex, in place (row pitch 2 nc reals == nc complex)
#pragma omp for collapse(2)
      for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) {
        T *r = p.data + ((size_t) i * n1 + j) * 2 * nc;
        for (int k = 0; k < n2; ++k) line[k] = C(r[k], 0);
        fft_line(line.data(), t2, -1);
        C *o = c + ((size_t) i * n1 + j) * nc;
        for (int k = 0; k < nc; ++k) o[k] = line[k];
      }
    }
#pragma omp for collapse(2)
    for (int i = 0; i < n0; ++i) for (int k = 0; k < nc; ++k) {
      for (int j = 0; j < n1; ++j) line[j] = c[((size_t) i * n1 + j) * nc + k];
      fft_line(line.data(), t1, sign);
      for (int j = 0; j < n1; ++j) c[((size_t) i * n1 + j) * nc + k] = line[j];
    }
#pragma omp for collapse(2)
    for (int j = 0; j < n1; ++j) for (int k = 0; k < nc; ++k) {
      for (int i = 0; i < n0; ++i) line[i] = c[((size_t) i * n1 + j) * nc + k];
      fft_line(line.data(), t0, sign);
      for (int i = 0; i < n0; ++i) c[((size_t) i * n1 + j) * nc + k] = line[i];
    }
    if (p.kind == 1) {  // rows last: Hermitian extension, backward transform, real parts
#pragma omp for collapse(2)
      for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) {
        const C *in = c + ((size_t) i * n1 + j) * nc;
        for (int k = 0; k < nc; ++k) line[k] = in[k];
        line[0] = C(line[0].real(), 0);
        if (n2 % 2 == 0) line[n2 / 2] = C(line[n2 / 2].real(), 0);
        for (int k = nc; k < n2; ++k) line[k] = std::conj(line[n2 - k]);
        fft_line(line.data(), t2, +1);
        T *r = p.data + ((size_t) i * n1 + j) * 2 * nc;
        for (int k = 0; k < n2; ++k) r[k] = line[k].real();
      }
    }
  }
}
}  // namespace

struct gx_shim_plan {
  int is_float;
  Plan<double> d;
  Plan<float> f;
};

static gx_shim_plan *make(int is_float, int n0, int n1, int n2, int kind, void *data) {
  gx_shim_plan *p = new gx_shim_plan();
  p->is_float = is_float;
  p->d = Plan<double>{n0, n1, n2, kind, (double *) data};
  p->f = Plan<float>{n0, n1, n2, kind, (float *) data};
  return p;
}

extern "C" {
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
fftw_plan fftw_plan_dft_r2c_3d(int a, int b, int c, double *in, fftw_complex *out, unsigned) {
  if ((void *) in != (void *) out) return nullptr;
  return make(0, a, b, c, 0, in);
}
fftw_plan fftw_plan_dft_c2r_3d(int a, int b, int c, fftw_complex *in, double *out, unsigned) {
  if ((void *) in != (void *) out) return nullptr;
  return make(0, a, b, c, 1, out);
}
fftw_plan fftw_plan_dft_3d(int a, int b, int c, fftw_complex *in, fftw_complex *out, int sign, unsigned) {
  if (in != out) return nullptr;
  return make(0, a, b, c, sign < 0 ? 2 : 3, in);
}
void fftw_execute(const fftw_plan p) { if (p->is_float) execute(p->f); else execute(p->d); }
void fftw_destroy_plan(fftw_plan p) { delete p; }
fftwf_plan fftwf_plan_dft_r2c_3d(int a, int b, int c, float *in, fftwf_complex *out, unsigned) {
  if ((void *) in != (void *) out) return nullptr;
  return make(1, a, b, c, 0, in);
}
fftwf_plan fftwf_plan_dft_c2r_3d(int a, int b, int c, fftwf_complex *in, float *out, unsigned) {
  if ((void *) in != (void *) out) return nullptr;
  return make(1, a, b, c, 1, out);
}
fftwf_plan fftwf_plan_dft_3d(int a, int b, int c, fftwf_complex *in, fftwf_complex *out, int sign, unsigned) {
  if (in != out) return nullptr;
  return make(1, a, b, c, sign < 0 ? 2 : 3, in);
}
void fftwf_execute(const fftwf_plan p) { fftw_execute(p); }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
}

// Power-spectrum host logic (product code; the test oracle has its own independent restatement).
//
// Mirrors cosmology::CAMB<T> of the reference (genetIC/src/cosmology/camb.hpp:173-332):
//   - log-log natural cubic spline through the transfer-function table (tools/numerics/interpolation.hpp:82-110,
//     GSL cspline semantics: zero second derivative at both ends, tridiagonal solve)
//   - sigma8 normalisation by a 50 000-step top-hat integral accumulated as k += dk (camb.hpp:300-328)
//   - linear growth factor with the reference's single-precision quirks (parameters.hpp:55-70)
// The spline tables go to the device once; kernels evaluate C(k) per mode (spectral.cuh).
#include "../../include/genetic_b200.h"
#include "spectrum_host.h"
#include <cmath>
#include <limits>
#include <vector>

struct gx_spectrum {
  int kind = 0;   // 0 CAMB, 1 power law
  int dtype = GX_F64;
  std::vector<double> lx, ly, c;          // spline of the applied species
  std::vector<double> nlx, nly, ncoef;    // spline of the species used for sigma8 (total)
  double amplitude = 0, ns = 0;
  double kfirst = 0, klast = 0;
  double *d_tab = nullptr;
};

namespace {

// Natural cubic spline coefficients c_i (half second derivatives) via LDL^T of the symmetric tridiagonal system.
void natural_spline(const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &c) {
  const size_t npts = x.size();
  c.assign(npts, 0.0);
  if (npts < 3) return;
  const size_t ns = npts - 2;
  std::vector<double> diag(ns), off(ns), rhs(ns);
  for (size_t i = 0; i < ns; ++i) {
    const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
    const double dy0 = y[i + 1] - y[i], dy1 = y[i + 2] - y[i + 1];
    const double g0 = (h0 != 0.0) ? 1.0 / h0 : 0.0, g1 = (h1 != 0.0) ? 1.0 / h1 : 0.0;
    off[i] = h1;
    diag[i] = 2.0 * (h1 + h0);
    rhs[i] = 3.0 * (dy1 * g1 - dy0 * g0);
  }
  if (ns == 1) { c[1] = rhs[0] / diag[0]; return; }
  std::vector<double> alpha(ns), gamma(ns), z(ns), w(ns);
  alpha[0] = diag[0];
  gamma[0] = off[0] / alpha[0];
  for (size_t i = 1; i + 1 < ns; ++i) {
    alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
    gamma[i] = off[i] / alpha[i];
  }
  alpha[ns - 1] = diag[ns - 1] - off[ns - 2] * gamma[ns - 2];
  z[0] = rhs[0];
  for (size_t i = 1; i < ns; ++i) z[i] = rhs[i] - gamma[i - 1] * z[i - 1];
  for (size_t i = 0; i < ns; ++i) w[i] = z[i] / alpha[i];
  c[ns] = w[ns - 1];
  for (size_t i = ns - 1; i-- > 0;) c[i + 1] = w[i] - gamma[i] * c[i + 2];
}

double spline_at(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &c, double v) {
  size_t lo = 0, hi = x.size() - 1;
  while (hi > lo + 1) { size_t mid = (hi + lo) / 2; if (x[mid] > v) hi = mid; else lo = mid; }
  const double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo], del = v - x[lo];
  const double b = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  const double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
  return y[lo] + del * (b + del * (c[lo] + del * d));
}

// LogInterpolator::operator() (interpolation.hpp:105-109)
double log_interp(const std::vector<double> &lx, const std::vector<double> &ly, const std::vector<double> &c, double k) {
  const double lg = std::log(k);
  if (lg < lx.front() || lg > lx.back()) return 0.0;
  return std::exp(spline_at(lx, ly, c, lg));
}

void build_log_spline(const double *k, const double *T, int nk, std::vector<double> &lx, std::vector<double> &ly,
                      std::vector<double> &c) {
  lx.clear(); ly.clear();
  for (int i = 0; i < nk; ++i) {
    const double a = std::log(k[i]), b = std::log(T[i]);
    if (std::isfinite(a) && std::isfinite(b)) { lx.push_back(a); ly.push_back(b); }
  }
  natural_spline(lx, ly, c);
}

// parameters.hpp:55-70, typed as the reference types it (FloatType F, powf on a, float exponent 4/7)
template <typename F> F growth(F Om, F Ol, F a) {
  F Hsq = Om / powf(a, 3.0) + (1. - Om - Ol) / a / a + Ol;
  F d = 2.5 * a * Om / std::pow(a, 3.0) / Hsq /
        (std::pow(Om / Hsq / a / a / a, 4.f / 7.f) - Ol / Hsq + (1. + 0.5 * Om / std::pow(a, 3.0) / Hsq) * (1. + 1. / 70. * Ol / Hsq));
  return d;
}

// camb.hpp:300-328
template <typename F> F tophat_sigma(const gx_spectrum &s, F ns, F radius) {
  F acc = 0., k, t;
  const F amp = 9. / 2. / M_PI / M_PI;
  const F kmax = std::min(s.klast, 200.0 / radius) * 0.999999;
  const F kmin = s.kfirst * 1.000001;
  const F dk = (kmax - kmin) / 50000.;
  for (k = kmin; k < kmax; k += dk) {
    t = log_interp(s.nlx, s.nly, s.ncoef, k);
    const auto kr = k * radius;
    acc += std::pow(k, ns + 2.) * ((std::sin(kr) - kr * std::cos(kr)) / (kr * kr * kr)) *
           ((std::sin(kr) - kr * std::cos(kr)) / (kr * kr * kr)) * t * t;
  }
  return std::sqrt(acc * amp * dk);
}

// camb.hpp:289-296
template <typename F> F overall_amplitude(const gx_spectrum &s, F Om, F Ol, F sigma8, F ns, F a) {
  const F g = growth<F>(Om, Ol, a);
  const F a_today = 1. / (1. + 0.f);
  const F gnorm = g / growth<F>(Om, Ol, a_today);
  const F s8pre = tophat_sigma<F>(s, ns, 8.);
  const F renorm = (sigma8 / s8pre) * gnorm;
  return renorm * renorm;
}

int upload_tables(gx_spectrum *s) {
  const size_t m = s->lx.size();
  if (m == 0) return 0;
  std::vector<double> tab(3 * m);
  for (size_t i = 0; i < m; ++i) { tab[i] = s->lx[i]; tab[m + i] = s->ly[i]; tab[2 * m + i] = s->c[i]; }
  GX_CUDA_CHECK(cudaMalloc(&s->d_tab, tab.size() * sizeof(double)));
  GX_CUDA_CHECK(cudaMemcpy(s->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

gx::SpecDev gx_spectrum_device_view(const gx_spectrum *s) {
  gx::SpecDev d;
  if (s->d_tab == nullptr && !s->lx.empty()) upload_tables(const_cast<gx_spectrum *>(s));  // lazy: host-only use needs no GPU
  d.kind = s->kind;
  d.m = (int) s->lx.size();
  d.tab = s->d_tab;
  d.amplitude = s->amplitude; d.ns = s->ns;
  d.amplitude_f = (float) s->amplitude; d.ns_f = (float) s->ns;
  return d;
}

double gx_spectrum_grid_norm(const gx_spectrum *s, double grid_length, int n) {
  const double n3 = (double) ((size_t) n * (size_t) n * (size_t) n);
  if (s->dtype == GX_F32) {
    const float L = (float) grid_length;
    const float kw = 2. * M_PI / L;
    const float norm = kw * kw * kw / std::pow(2.f * M_PI, 3.f);
    const float fftn = (float) ((size_t) n * (size_t) n * (size_t) n);
    return (double) (norm * fftn);
  }
  const double kw = 2. * M_PI / grid_length;
  const double norm = kw * kw * kw / std::pow(2.f * M_PI, 3.f);
  return norm * n3;
}

extern "C" {

int gx_spectrum_create_camb(const double *k, const double *T, const double *T_norm, int nk, double OmegaM0,
                            double OmegaLambda0, double sigma8, double ns, double scalefactor, int dtype,
                            gx_spectrum **out) {
  if (nk < 2) return gx::fail("gx_spectrum_create_camb: need at least two table rows");
  if (dtype != GX_F32 && dtype != GX_F64) return gx::fail("bad dtype");
  gx_spectrum *s = new gx_spectrum();
  s->kind = 0; s->dtype = dtype; s->ns = ns;
  s->kfirst = k[0]; s->klast = k[nk - 1];
  build_log_spline(k, T, nk, s->lx, s->ly, s->c);
  build_log_spline(k, T_norm, nk, s->nlx, s->nly, s->ncoef);
  if (s->lx.size() < 2 || s->nlx.size() < 2) {
    delete s;
    return gx::fail("Could not find enough valid values in the transfer function. This may mean that one of the "
                    "transfer function columns is full of zeros.");
  }
  if (dtype == GX_F32)
    s->amplitude = overall_amplitude<float>(*s, (float) OmegaM0, (float) OmegaLambda0, (float) sigma8, (float) ns,
                                            (float) scalefactor);
  else
    s->amplitude = overall_amplitude<double>(*s, OmegaM0, OmegaLambda0, sigma8, ns, scalefactor);
  *out = s;
  return 0;
}

int gx_spectrum_create_powerlaw(double amplitude, double ns, int dtype, gx_spectrum **out) {
  gx_spectrum *s = new gx_spectrum();
  s->kind = 1; s->dtype = dtype; s->amplitude = amplitude; s->ns = ns;
  *out = s;
  return 0;
}

int gx_spectrum_destroy(gx_spectrum *s) {
  if (!s) return 0;
  if (s->d_tab) cudaFree(s->d_tab);
  delete s;
  return 0;
}

double gx_spectrum_amplitude(const gx_spectrum *s) { return s->amplitude; }

double gx_spectrum_eval(const gx_spectrum *s, double k) {
  if (s->kind == 1) return k == 0 ? 0.0 : s->amplitude * std::pow(k, s->ns);
  const double lt = (k != 0) ? log_interp(s->lx, s->ly, s->c, k) : 0.0;
  return s->amplitude * std::pow(k, s->ns) * lt * lt;
}

}  // extern "C"

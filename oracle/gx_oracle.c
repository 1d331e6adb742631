/* TEST INFRASTRUCTURE -- CPU oracle, not product code.
 *
 * Plain-C restatement of the serial pieces of genetIC's white-noise path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * What it follows (paths relative to /root/reference/genetIC/src):
 *   - simulation/field/randomfieldgenerator.hpp:43-45   gsl_rng_ranlxs2, gsl_rng_set
 *   - simulation/field/randomfieldgenerator.hpp:144-171 drawOneFourierMode (incl. the -k write quirk)
 *   - simulation/field/randomfieldgenerator.hpp:181-197 real-space draw (before the r2c)
 *   - simulation/field/randomfieldgenerator.hpp:203-260 square-shell Fourier draw, serial and per-shell seeded
 *   - tools/numerics/fourier.hpp:300-319,439-453        setFourierCoefficient slot + conjugation rule
 * GSL itself is an un-vendored system dependency of the reference (GSL 2.7.x on the reference's
 * Docker/CI images); ranlxs2 (Luescher's RANLUX v3, 48-bit words, luxury p=397) and the 128-level
 * ziggurat (J. Voss' gausszig.c) are restated here from their published algorithms. Pinned by the
 * reference's own goldens: tests/test_02k_chisquare_onelevel/reference.txt (chi^2 to 10 digits),
 * tests/test_13_parallel_seed and tests/test_05a_seed_old (float32 particles, bit-exact) -- see
 * tests/test_oracle_golden.py.
 */
#include <tgmath.h> /* type-generic: pow(float,float)->powf etc., mirroring the C++ std:: overloads the reference uses */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "zig_tables.h"

typedef struct {
  double xdbl[12], ydbl[12];
  double carry;
  float xflt[24];
  unsigned int ir, jr, is, is_old, pr;
} orc_ranlxs_t;

static const int nxt[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0};
static const double one_bit = 1.0 / 281474976710656.0; /* 2^-48 */
static const double sone_bit = 1.0 / 16777216.0;       /* 2^-24 */
static const double sbase = 16777216.0;                /* 2^24 */
static const double shiftc = 268435456.0;              /* 2^28 */

size_t orc_state_size(void) { return sizeof(orc_ranlxs_t); }

/* gsl_rng_set for ranlxs2: 31-bit seed expanded by a 31-bit LFSR into 12 48-bit words */
void orc_ranlxs2_set(orc_ranlxs_t *st, unsigned long s) {
  int xbit[31];
  if (s == 0) s = 1;
  long i = (long) (s & 0x7fffffffUL);
  for (int k = 0; k < 31; ++k) { xbit[k] = (int) (i % 2); i /= 2; }
  int ibit = 0, jbit = 18;
  for (int k = 0; k < 12; ++k) {
    double x = 0;
    for (int m = 1; m <= 48; ++m) {
      double y = (double) xbit[ibit];
      x += x + y;
      xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
      ibit = (ibit + 1) % 31;
      jbit = (jbit + 1) % 31;
    }
    st->xdbl[k] = one_bit * x;
  }
  st->carry = 0;
  st->ir = 0; st->jr = 7; st->is = 23; st->is_old = 0; st->pr = 397;
}

/* One subtract-with-borrow word step: x[ir] <- x[jr] - x[ir] - carry (mod 1), base 2^48 */
static inline double swb_step(double *xdbl, unsigned ir, unsigned jr, double *carry) {
  double y = xdbl[jr] - xdbl[ir];
  y -= *carry;
  if (y < 0) { *carry = one_bit; y += 1; } else { *carry = 0; }
  xdbl[ir] = y;
  return y;
}

static void increment_state(orc_ranlxs_t *st) {
  double *xdbl = st->xdbl, *ydbl = st->ydbl;
  double carry = st->carry;
  unsigned ir = st->ir, jr = st->jr;
  int k;
  /* pr word steps in total; the last (pr % 12 ... ) steps also record y + 2^28 (the rounding trick
     that later splits each 48-bit word into two 24-bit halves). Every word gets a ydbl in the end,
     so we simply do all steps and fill ydbl afterwards -- same values. */
  for (k = 0; k < (int) st->pr; ++k) {
    swb_step(xdbl, ir, jr, &carry);
    ir = nxt[ir]; jr = nxt[jr];
  }
  for (k = 0; k < 12; ++k) ydbl[k] = xdbl[k] + shiftc;
  for (int kk = 0, m = 0; kk < 12; ++kk) {
    double x = xdbl[kk];
    double y2 = ydbl[kk] - shiftc;
    if (y2 > x) y2 -= sone_bit;
    double y1 = (x - y2) * sbase;
    st->xflt[m++] = (float) y1;
    st->xflt[m++] = (float) y2;
  }
  st->ir = ir; st->jr = jr; st->is = 2 * ir; st->is_old = 2 * ir; st->carry = carry;
}

double orc_ranlxs_get_double(orc_ranlxs_t *st) {
  unsigned is = (st->is + 1) % 24;
  st->is = is;
  if (is == st->is_old) increment_state(st);
  return st->xflt[st->is];
}

static inline unsigned long rng_get(orc_ranlxs_t *st) {
  return (unsigned long) (orc_ranlxs_get_double(st) * 16777216.0);
}

/* gsl_ran_gaussian_ziggurat(r, 1.0) for a generator with range 2^24-1 (two raw draws per attempt) */
double orc_gaussian_ziggurat(orc_ranlxs_t *st) {
  unsigned long i, j;
  int sign;
  double x, y;
  while (1) {
    unsigned long k1 = rng_get(st);
    unsigned long k2 = rng_get(st);
    i = (k1 & 0xFF);
    j = (k2 & 0x00FFFFFF);
    sign = (i & 0x80) ? +1 : -1;
    i &= 0x7f;
    x = j * gx_zig_wtab[i];
    if (j < gx_zig_ktab[i]) break;
    if (i < 127) {
      double y0 = gx_zig_ytab[i], y1 = gx_zig_ytab[i + 1];
      double U1 = orc_ranlxs_get_double(st);
      y = y1 + (y0 - y1) * U1;
    } else {
      double U1 = 1.0 - orc_ranlxs_get_double(st);
      double U2 = orc_ranlxs_get_double(st);
      x = GX_ZIG_R - log(U1) / GX_ZIG_R;
      y = exp(-GX_ZIG_R * (x - 0.5 * GX_ZIG_R)) * U2;
    }
    if (y < exp(-0.5 * x * x)) break;
  }
  return sign * 1.0 * x;
}

/* n raw 24-bit outputs (as integers) from a freshly seeded stream -- used to test the device generator */
void orc_raw_stream(unsigned long seed, long n, unsigned int *out) {
  orc_ranlxs_t st;
  orc_ranlxs2_set(&st, seed);
  for (long i = 0; i < n; ++i) out[i] = (unsigned int) rng_get(&st);
}

/* n gaussians from the stream in *st (continues the stream; caller seeds) */
void orc_gaussians(orc_ranlxs_t *st, long n, double *out) {
  for (long i = 0; i < n; ++i) out[i] = orc_gaussian_ziggurat(st);
}

/* randomfieldgenerator.hpp:181-192 -- N^3 gaussians in cell-index order, cast to the field type */
void orc_draw_real(orc_ranlxs_t *st, long n3, int is_float, double *out) {
  for (long i = 0; i < n3; ++i) {
    double g = orc_gaussian_ziggurat(st);
    out[i] = is_float ? (double) (float) g : g;
  }
}

/* fourier.hpp:300-319 + 439-453 */
static inline void set_coeff(double *data, int N, int kx, int ky, int kz, double re, double im) {
  int conj = 0;
  if (kz < 0) { conj = 1; kx = -kx; ky = -ky; kz = -kz; }
  if (ky < 0) ky += N;
  if (kx < 0) kx += N;
  size_t cs = (size_t) (N / 2 + 1);
  size_t idx = 2 * ((size_t) kz + cs * (size_t) ky + cs * (size_t) N * (size_t) kx);
  if (conj) im = -im;
  data[idx] = re;
  data[idx + 1] = im;
}

/* randomfieldgenerator.hpp:144-171 */
static inline void draw_one_mode(double *data, int N, int k1, int k2, int k3, int is_float, int reverse,
                                 orc_ranlxs_t *st) {
  double a, b;
  if (is_float) {
    float sigma = (float) (1.0 / sqrt(2.0));
    a = (double) (float) (sigma * orc_gaussian_ziggurat(st));
    b = (double) (float) (sigma * orc_gaussian_ziggurat(st));
  } else {
    double sigma = 1.0 / sqrt(2.0);
    a = sigma * orc_gaussian_ziggurat(st);
    b = sigma * orc_gaussian_ziggurat(st);
  }
  if (reverse) set_coeff(data, N, k1, k2, k3, b, a);
  else set_coeff(data, N, k1, k2, k3, a, b);
  int nyq = N / 2;
  if (k1 == 0 || k1 == nyq || k2 == 0 || k2 == nyq || k3 == 0 || k3 == nyq) {
    if (reverse) set_coeff(data, N, -k1, -k2, -k3, b, -a);
    else set_coeff(data, N, -k1, -k2, -k3, a, -b);
  }
}

/* randomfieldgenerator.hpp:203-260. data: 2*N*N*(N/2+1) doubles, zero-filled here.
   parallel != 0: one fresh stream per shell seeded base_seed+ks (st untouched);
   otherwise the single stream *st is consumed (and left advanced, as the reference leaves it). */
void orc_draw_fourier(orc_ranlxs_t *st, unsigned long base_seed, int N, int parallel, int reverse, int is_float,
                      double *data) {
  size_t len = 2 * (size_t) N * (size_t) N * (size_t) (N / 2 + 1);
  memset(data, 0, len * sizeof(double));
  for (int ks = 0; ks < N / 2; ks++) {
    orc_ranlxs_t local;
    orc_ranlxs_t *use = st;
    if (parallel) { orc_ranlxs2_set(&local, base_seed + (unsigned long) ks); use = &local; }
    for (int k1 = -ks; k1 < ks; k1++) {
      for (int k2 = -ks; k2 < ks; k2++) {
        draw_one_mode(data, N, ks, k1, k2, is_float, reverse, use);
        draw_one_mode(data, N, k1, ks, k2, is_float, reverse, use);
        draw_one_mode(data, N, k1, k2, ks, is_float, reverse, use);
      }
    }
  }
}

/* Timing helper for bench.py's cpu_baseline: shells [ks0, ks1) of the parallel Fourier draw, OpenMP over
   shells with schedule(dynamic) exactly as randomfieldgenerator.hpp:226-243. */
void orc_draw_fourier_parallel_shells(unsigned long base_seed, int N, int ks0, int ks1, int is_float, double *data) {
#pragma omp parallel for schedule(dynamic)
  for (int ks = ks0; ks < ks1; ks++) {
    orc_ranlxs_t local;
    orc_ranlxs2_set(&local, base_seed + (unsigned long) ks);
    for (int k1 = -ks; k1 < ks; k1++) {
      for (int k2 = -ks; k2 < ks; k2++) {
        draw_one_mode(data, N, ks, k1, k2, is_float, 0, &local);
        draw_one_mode(data, N, k1, ks, k2, is_float, 0, &local);
        draw_one_mode(data, N, k1, k2, ks, is_float, 0, &local);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Natural cubic spline, restating GSL 2.x interpolation/cspline.c (cspline_init, cspline_eval),
 * linalg/tridiag.c (solve_tridiag) and interpolation/bsearch.h, as used through
 * tools/numerics/interpolation.hpp:46-54,72-76,82-110 (LogInterpolator).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int n;          /* knots kept after dropping non-finite logs */
  double *lx;     /* ln k */
  double *ly[3];  /* ln T for dm(1), baryon(2), all(6) */
  double *c[3];   /* spline second-derivative coefficients */
  int nvalid[3];
  double *lxs[3]; /* per-species knot abscissae (knots with non-finite ln T are dropped per species) */
  double k_first, k_last;
} orc_camb_t;

static void solve_tridiag(const double *diag, const double *offdiag, const double *b, double *x, size_t N) {
  double *gamma = malloc(N * sizeof(double)), *alpha = malloc(N * sizeof(double));
  double *c = malloc(N * sizeof(double)), *z = malloc(N * sizeof(double));
  size_t i, j;
  alpha[0] = diag[0];
  gamma[0] = offdiag[0] / alpha[0];
  for (i = 1; i + 1 < N; i++) {
    alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
    gamma[i] = offdiag[i] / alpha[i];
  }
  if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
  z[0] = b[0];
  for (i = 1; i < N; i++) z[i] = b[i] - gamma[i - 1] * z[i - 1];
  for (i = 0; i < N; i++) c[i] = z[i] / alpha[i];
  x[N - 1] = c[N - 1];
  if (N >= 2)
    for (i = N - 2, j = 0; j <= N - 2; j++, i--) x[i] = c[i] - gamma[i] * x[i + 1];
  free(gamma); free(alpha); free(c); free(z);
}

static void cspline_init(const double *xa, const double *ya, int size, double *c) {
  size_t max_index = size - 1, sys_size = max_index - 1;
  c[0] = 0.0; c[max_index] = 0.0;
  if (sys_size == 0) return;
  double *g = malloc(sys_size * sizeof(double)), *diag = malloc(sys_size * sizeof(double));
  double *offdiag = malloc(sys_size * sizeof(double));
  for (size_t i = 0; i < sys_size; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double ydiff_i = ya[i + 1] - ya[i], ydiff_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    offdiag[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
  }
  if (sys_size == 1) c[1] = g[0] / diag[0];
  else solve_tridiag(diag, offdiag, g, c + 1, sys_size);
  free(g); free(diag); free(offdiag);
}

static inline double cspline_eval(const double *xa, const double *ya, const double *c, int size, double x) {
  size_t ilo = 0, ihi = size - 1;
  while (ihi > ilo + 1) { size_t i = (ihi + ilo) / 2; if (xa[i] > x) ihi = i; else ilo = i; }
  const double x_hi = xa[ilo + 1], x_lo = xa[ilo], dx = x_hi - x_lo;
  const double y_lo = ya[ilo], y_hi = ya[ilo + 1], dy = y_hi - y_lo, delx = x - x_lo;
  const double c_i = c[ilo], c_ip1 = c[ilo + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}

/* camb.hpp:224-268 (rows with k>0; transfer columns normalised by input[1]) + interpolation.hpp:88-103.
   k[n], and T[3][n] for species slots 0=dm(col 1), 1=baryon(col 2), 2=all(col 6), already divided by input[1]. */
orc_camb_t *orc_camb_create(const double *k, const double *Tdm, const double *Tb, const double *Tall, int n) {
  orc_camb_t *c = calloc(1, sizeof(orc_camb_t));
  const double *T[3] = {Tdm, Tb, Tall};
  c->n = n;
  c->k_first = k[0]; c->k_last = k[n - 1];
  for (int s = 0; s < 3; ++s) {
    c->lxs[s] = malloc(n * sizeof(double)); c->ly[s] = malloc(n * sizeof(double)); c->c[s] = calloc(n, sizeof(double));
    int m = 0;
    for (int i = 0; i < n; ++i) {
      double lx = log(k[i]), ly = log(T[s][i]);
      if (isfinite(lx) && isfinite(ly)) { c->lxs[s][m] = lx; c->ly[s][m] = ly; ++m; }
    }
    c->nvalid[s] = m;
    if (m >= 2) cspline_init(c->lxs[s], c->ly[s], m, c->c[s]);
  }
  return c;
}

void orc_camb_destroy(orc_camb_t *c) {
  for (int s = 0; s < 3; ++s) { free(c->lxs[s]); free(c->ly[s]); free(c->c[s]); }
  free(c);
}

/* interpolation.hpp:105-109 and 72-76 */
static inline double orc_loginterp_eval(const orc_camb_t *c, int s, double x) {
  double logx = log(x);
  int m = c->nvalid[s];
  double minx = c->lxs[s][0], maxx = c->lxs[s][m - 1];
  if (logx < minx || logx > maxx) return 0.0;
  return exp(cspline_eval(c->lxs[s], c->ly[s], c->c[s], m, logx));
}

double orc_camb_transfer(const orc_camb_t *c, int species, double k) { return orc_loginterp_eval(c, species, k); }

int orc_camb_spline_export(const orc_camb_t *c, int s, double *lx, double *ly, double *cc) {
  int m = c->nvalid[s];
  if (lx) { memcpy(lx, c->lxs[s], m * sizeof(double)); memcpy(ly, c->ly[s], m * sizeof(double)); memcpy(cc, c->c[s], m * sizeof(double)); }
  return m;
}

/* fourier.hpp:110-151: is storage slot (ix,iy,iz) reached by the half-space mode iteration?
   The iteration visits kx in [0,N/2], ky in [ky_lower,N/2], kz in [kz_lower,N/2] and maps kz<0 to the
   conjugate slot. Never reached: on planes iz in {0,N/2}: ix>N/2, or ix in {0,N/2} and iy>N/2. */
int orc_slot_is_independent(int N, int ix, int iy, int iz) {
  int h = N / 2;
  if (iz != 0 && iz != h) return 1;
  if (ix > h) return 0;
  if ((ix == 0 || ix == h) && iy > h) return 0;
  return 1;
}

#define SPECIES_COL 2 /* particle::species::all -> CAMB column 6, camb.hpp:181-185,301 */
#define REAL double
#define SFX _d
#include "cosmo_impl.inc"
#undef REAL
#undef SFX
#define REAL float
#define SFX _f
#include "cosmo_impl.inc"
#undef REAL
#undef SFX

/* ---- exported for oracle/shims/gsl_shim.cpp (the reference itself compiled against stand-in headers) ------------------- */
void orc_cspline_init(const double *xa, const double *ya, int size, double *c) { cspline_init(xa, ya, size, c); }
double orc_cspline_eval(const double *xa, const double *ya, const double *c, int size, double x) {
  return cspline_eval(xa, ya, c, size, x);
}

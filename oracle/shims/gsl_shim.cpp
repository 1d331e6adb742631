// TEST INFRASTRUCTURE: the GSL entry points the reference calls, forwarded to the restated algorithms in
// oracle/gx_oracle.c (ranlxs2, 128-level ziggurat, natural cubic spline). See oracle/shims/README.md.
#include "gsl/gsl_randist.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_spline.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {
size_t orc_state_size(void);
void orc_ranlxs2_set(void *st, unsigned long s);
double orc_ranlxs_get_double(void *st);
double orc_gaussian_ziggurat(void *st);
void orc_cspline_init(const double *xa, const double *ya, int size, double *c);
double orc_cspline_eval(const double *xa, const double *ya, const double *c, int size, double x);
}

struct gsl_rng_type { int id; };
struct gsl_rng { void *state; };
static const gsl_rng_type ranlxs2_type = {2};
extern "C" const gsl_rng_type *gsl_rng_ranlxs2 = &ranlxs2_type;

extern "C" gsl_rng *gsl_rng_alloc(const gsl_rng_type *) {
  gsl_rng *r = new gsl_rng();
  r->state = std::calloc(1, orc_state_size());
  orc_ranlxs2_set(r->state, 0);  // gsl_rng_alloc seeds with the default seed (0 -> 1 for ranlxs)
  return r;
}
extern "C" void gsl_rng_set(const gsl_rng *r, unsigned long seed) { orc_ranlxs2_set(r->state, seed); }
extern "C" void gsl_rng_free(gsl_rng *r) { if (r) { std::free(r->state); delete r; } }
extern "C" double gsl_rng_uniform(const gsl_rng *r) { return orc_ranlxs_get_double(r->state); }
extern "C" double gsl_ran_gaussian_ziggurat(const gsl_rng *r, double sigma) { return sigma * orc_gaussian_ziggurat(r->state); }

struct gsl_interp_type { int id; };
struct gsl_interp_accel { int unused; };
struct gsl_spline { std::vector<double> x, y, c; };
static const gsl_interp_type cspline_type = {1};
extern "C" const gsl_interp_type *gsl_interp_cspline = &cspline_type;
extern "C" gsl_interp_accel *gsl_interp_accel_alloc(void) { return new gsl_interp_accel(); }
extern "C" void gsl_interp_accel_free(gsl_interp_accel *a) { delete a; }
extern "C" gsl_spline *gsl_spline_alloc(const gsl_interp_type *, size_t size) {
  gsl_spline *s = new gsl_spline();
  s->x.resize(size); s->y.resize(size); s->c.resize(size);
  return s;
}
extern "C" int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size) {
  s->x.assign(xa, xa + size); s->y.assign(ya, ya + size); s->c.assign(size, 0.0);
  orc_cspline_init(s->x.data(), s->y.data(), (int) size, s->c.data());
  return 0;
}
extern "C" double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *) {
  if (x < s->x.front() || x > s->x.back()) {  // GSL's default handler aborts on a domain error
    std::fprintf(stderr, "gsl shim: interpolation error (x = %g outside [%g, %g])\n", x, s->x.front(), s->x.back());
    std::abort();
  }
  return orc_cspline_eval(s->x.data(), s->y.data(), s->c.data(), (int) s->x.size(), x);
}
extern "C" void gsl_spline_free(gsl_spline *s) { delete s; }

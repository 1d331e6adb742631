/* TEST INFRASTRUCTURE: stand-in for <gsl/gsl_rng.h> (oracle/shims/README.md): ranlxs2 through oracle/gx_oracle.c. */
#ifndef GX_SHIM_GSL_RNG_H
#define GX_SHIM_GSL_RNG_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gsl_rng_type gsl_rng_type;
typedef struct gsl_rng gsl_rng;
extern const gsl_rng_type *gsl_rng_ranlxs2;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
#ifdef __cplusplus
}
#endif
#endif

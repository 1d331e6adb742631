/* TEST INFRASTRUCTURE: stand-in for <gsl/gsl_randist.h> (oracle/shims/README.md). */
#ifndef GX_SHIM_GSL_RANDIST_H
#define GX_SHIM_GSL_RANDIST_H
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_gaussian_ziggurat(const gsl_rng *r, double sigma);
#ifdef __cplusplus
}
#endif
#endif

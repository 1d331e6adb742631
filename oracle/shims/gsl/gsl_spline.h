/* TEST INFRASTRUCTURE: stand-in for <gsl/gsl_spline.h> (oracle/shims/README.md): natural cubic spline of oracle/gx_oracle.c. */
#ifndef GX_SHIM_GSL_SPLINE_H
#define GX_SHIM_GSL_SPLINE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gsl_interp_type gsl_interp_type;
typedef struct gsl_interp_accel gsl_interp_accel;
typedef struct gsl_spline gsl_spline;
extern const gsl_interp_type *gsl_interp_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *spline, const double *xa, const double *ya, size_t size);
double gsl_spline_eval(const gsl_spline *spline, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *spline);
#ifdef __cplusplus
}
#endif
#endif

/* TEST INFRASTRUCTURE: stand-in for <gsl/gsl_errno.h> (nothing of it is used by the reference beyond the include). */
#ifndef GX_SHIM_GSL_ERRNO_H
#define GX_SHIM_GSL_ERRNO_H
#endif

/* TEST INFRASTRUCTURE: minimal stand-in for <fftw3.h> (see oracle/shims/README.md). Only what genetIC's
   tools/numerics/fourier.hpp uses. Transforms are unnormalised and in place on FFTW's padded real layout. */
#ifndef GX_SHIM_FFTW3_H
#define GX_SHIM_FFTW3_H
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef float fftwf_complex[2];
typedef struct gx_shim_plan *fftw_plan;
typedef struct gx_shim_plan *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int n);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex *in, double *out, unsigned flags);
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
fftwf_plan fftwf_plan_dft_r2c_3d(int n0, int n1, int n2, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_3d(int n0, int n1, int n2, fftwf_complex *in, float *out, unsigned flags);
fftwf_plan fftwf_plan_dft_3d(int n0, int n1, int n2, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif

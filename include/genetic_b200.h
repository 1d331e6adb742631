/* genetic_b200 -- C ABI of the B200-native genetIC hot path (Gaussian random field generation +
 * Zeldovich displacement).  extern "C", plain pointers and sizes, no C++/torch types.
 *
 * The reference (pynbody/genetIC) has no FFI: its seam is the C++ class API of fields::Field /
 * tools::numerics::fourier::FieldFourierManager / fields::RandomFieldGenerator / cosmology::CAMB /
 * particle::ZeldovichParticleGenerator.  Each entry point below names the reference member function(s)
 * it replaces (paths relative to genetIC/src/).  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *  - every function returns 0 on success; otherwise non-zero and gx_last_error() describes the failure
 *    (the C++ host layer turns that into std::runtime_error, the reference's error behaviour,
 *    tools/parser.hpp:202-207).
 *  - `dtype`: GX_F32 (the reference's `-f` build, main.cpp:14-20) or GX_F64.
 *  - a *field* is a device buffer of gx_field_bytes(n, dtype) bytes: n*n rows of (n/2+1) complex =
 *    n+2 reals, z fastest (grid.hpp:610-626, fourier.hpp:300-319, 471-476).  Fourier space uses every
 *    slot; real space uses the first n reals of each row (the device keeps the padded pitch; the
 *    reference's unpadded N^3 real-space layout exists only on the host side of gx_field_upload /
 *    gx_field_download).
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 *    asynchronous with respect to the host unless documented otherwise.
 *  - n must be even and in [4, 4096]. Powers of two take the tuned passes; other even sizes the mixed-radix passes
 *    (fft_generic.cu) -- every element-wise entry point and the real-space draw order accept them, while the Fourier-space
 *    draw orders, gx_unigrid_ic and the gx_slab_* entry points need a power of two and fail with a message otherwise.
 */
#ifndef GENETIC_B200_H
#define GENETIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_F32 0
#define GX_F64 1

/* host layouts for upload/download */
#define GX_LAYOUT_REAL 0    /* n^3 reals, unpadded, C order [x][y][z]      (Field::getDataVector() in real space) */
#define GX_LAYOUT_FOURIER 1 /* 2*n*n*(n/2+1) reals, the padded half-complex (Field::getDataVector() in Fourier space) */

/* white-noise draw orders (randomfieldgenerator.hpp) */
#define GX_NOISE_REAL 0             /* drawRandomForSpecifiedGrid :181-197 (+ r2c, Re(DC)=0) */
#define GX_NOISE_FOURIER_SERIAL 1   /* drawRandomForSpecifiedGridFourier :246-256 */
#define GX_NOISE_FOURIER_PARALLEL 2 /* drawRandomForSpecifiedGridFourier :226-243 (stream per shell, seed+ks) */

/* ---- library ---------------------------------------------------------------------------------- */
const char *gx_last_error(void);
int gx_version(void);
int gx_device_count(int *count);
int gx_set_device(int device);           /* fourier::initialise() analogue, fourier.hpp:36-81 */
int gx_synchronize(void *stream);
int gx_workspace_release(void);          /* frees the grow-only scratch the library keeps between calls (noise workspace, |k|^2 table, cross-grid
                                            tables) and the idle field buffers of the gx_malloc pool */
unsigned long long gx_launch_count(void); /* kernels launched by this library so far (bench bookkeeping) */

/* ---- memory ----------------------------------------------------------------------------------- */
size_t gx_field_bytes(int n, int dtype); /* FieldFourierManager::getRequiredDataSize, fourier.hpp:471-476 */
/* Field buffers (the reference: std::vector inside fields::Field, copied freely). Freed blocks up to GX_POOL_MAX_BLOCK bytes
   (default 1.5 GiB) are pooled by exact size, at most GX_POOL_BYTES (default 8 GiB) in total, and handed out again once the
   work issued before the free (legacy default stream and blocking streams) has completed; GX_POOL=0 maps and unmaps every
   block (cudaMalloc / cudaFree). A buffer used on a non-blocking stream is synchronised by its owner before gx_free. */
int gx_malloc(void **ptr, size_t bytes);
int gx_free(void *ptr);
int gx_memset_zero(void *ptr, size_t bytes, void *stream);
int gx_host_alloc_pinned(void **ptr, size_t bytes);
/* Peer memory between the ranks of one node (one process per GPU; no reference counterpart -- the reference is a single
   process): CUDA IPC handle (64 bytes) of an allocation made by gx_malloc, opened by the other ranks as a device pointer
   their kernels can store through (NVLink). Used by gx_slab_noise_staged and gx_slab_fft_x_mirror_push. */
int gx_ipc_get_handle(void *ptr, void *handle_out_64_bytes);
int gx_ipc_open(const void *handle_64_bytes, void **ptr_out);
int gx_ipc_close(void *ptr);
int gx_host_free_pinned(void *ptr);
int gx_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int gx_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int gx_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
/* Field(grid, dataVector, fourier) construction / getDataVector() read-back (field.hpp:230-249,271-278).
   GX_LAYOUT_REAL repacks between the host's unpadded n^3 and the device's padded pitch
   (replaces padForFFTWRealTransform / unpadAfterFFTWRealTransform, fourier.hpp:368-405). */
int gx_field_upload(void *dev_field, const void *host, int n, int dtype, int layout, void *stream);
int gx_field_download(void *host, const void *dev_field, int n, int dtype, int layout, void *stream);

/* ---- FFT (FieldFourierManager<T,T>::performTransform, fourier.hpp:479-518) ----------------------- */
int gx_fft_forward(void *field, int n, int dtype, void *stream); /* toFourier: r2c, / n^1.5 */
int gx_fft_inverse(void *field, int n, int dtype, void *stream); /* toReal: mirror, c2r, / n^1.5 */
int gx_mirror(void *field, int n, int dtype, void *stream);      /* ensureFourierModesAreMirrored :331-357 */

/* ---- white noise (RandomFieldGenerator::draw, randomfieldgenerator.hpp:122-139) ------------------ */
/* Draws one level. `stream_offset_io` (may be NULL): for the two serial modes the reference draws all
   levels from ONE gsl stream; pass the number of raw 24-bit outputs already consumed (0 for the first
   level) and receive the updated count. The result is in Fourier space for every mode. */
int gx_noise_draw(void *field, int n, int dtype, unsigned long seed, int mode, int reverse_order,
                  unsigned long long *stream_offset_io, void *stream);
/* Test hooks for the generator: n raw 24-bit outputs / n gaussians of gsl_rng_ranlxs2 seeded with `seed`,
   starting at raw offset `offset`, computed on the device and written to host arrays. */
int gx_rng_raw(unsigned long seed, unsigned long long offset, long n, uint32_t *host_out);
int gx_rng_gaussians(unsigned long seed, long n, double *host_out, unsigned long long *raw_consumed);

/* ---- power spectrum (cosmology::CAMB, camb.hpp:173-332; PowerLawPowerSpectrum :150-168) ---------- */
typedef struct gx_spectrum gx_spectrum;
/* k[nk] > 0 and T[nk] = one CAMB column already divided by the file's first CDM entry (camb.hpp:254-262);
   calculates the spline (interpolation.hpp:46-54,88-103) and the sigma8/growth normalisation (:289-328)
   with T_norm[nk] = the `total` column (species::all), and uploads the tables. */
int gx_spectrum_create_camb(const double *k, const double *T, const double *T_norm, int nk, double OmegaM0,
                            double OmegaLambda0, double sigma8, double ns, double scalefactor, int dtype,
                            gx_spectrum **out);
int gx_spectrum_create_powerlaw(double amplitude, double ns, int dtype, gx_spectrum **out);
int gx_spectrum_destroy(gx_spectrum *s);
double gx_spectrum_amplitude(const gx_spectrum *s);
double gx_spectrum_eval(const gx_spectrum *s, double k); /* CAMB::operator()(k), host evaluation */
/* cosmology::dumpPowerSpectrum (cosmology/parameters.hpp:112-205), the binning reduction of `dump_ps` on the device: both
   fields are mirrored in place (ensureFourierModesAreMirrored, :121-122), every (ix, iy, iz) in [-n/2, n/2]^3 of the reference's
   loop is accounted for through the multiplicity of its storage slot, and host_bins[4][nbins] receives per logarithmic k bin:
   the count, the sum of k, the sum of |delta_k|^2 and the sum of Re C(k) (cov_field may be NULL: zeros). The caller divides
   by the count and by the grid's power-spectrum normalisation and writes the text file. */
int gx_power_spectrum_bins(void *field, void *cov_field, int n, int dtype, double grid_length, int nbins, double *host_bins,
                           void *stream);
/* getPowerSpectrumForGridUncached (camb.hpp:101-119): materialise C(k) = P(k) * n^3 / L^3 as a Fourier field */
int gx_covariance_field(void *out_field, int n, int dtype, double grid_length, const gx_spectrum *s, void *stream);
/* Field::applyTransferFunction (field.hpp:394-412) with the covariance evaluated in-kernel (nothing materialised) */
int gx_apply_transfer(void *field, int n, int dtype, double grid_length, const gx_spectrum *s, double power,
                      void *stream);
/* Field::applyTransferFunction with an explicit covariance field */
int gx_apply_transfer_field(void *field, const void *cov_field, int n, int dtype, double power, void *stream);

/* ---- filters (filters::*, filter.hpp:107-317; Field::applyFilter, field.hpp:741-746) -------------- */
#define GX_FILTER_CONST 0          /* value                                       (filter.hpp:60-105)  */
#define GX_FILTER_FERMI_LP 1       /* 1/(1+exp((k-kcut)/T)), T = kcut*param       (filter.hpp:209-239) */
#define GX_FILTER_FERMI_HP_COV 2   /* sqrt(1 - lp^2)                              (filter.hpp:246-275) */
#define GX_FILTER_FERMI_HP 3       /* 1 - lp                                      (filter.hpp:282-317) */
typedef struct {
  int kind;
  double kcut;  /* or the constant */
  double temperature;
  int power;    /* factor raised to this integer power (>=1); f*f products fold to power 2 */
} gx_filter_factor;
int gx_apply_filter(void *field, int n, int dtype, double grid_length, const gx_filter_factor *factors,
                    int nfactors, void *stream);

/* ---- element-wise and reductions ------------------------------------------------------------------- */
int gx_unit_variance(void *field, int n, int dtype, void *stream); /* enforceUnitVariance, multilevelfield.hpp:498-510 */
/* ICGenerator::reverseSmallK (ic.hpp:1690-1708): the sign of every mode with 0 < k^2 < kmax^2 reversed (Fourier space) */
int gx_reverse_small_k(void *field, int n, int dtype, double grid_length, double kmax, void *stream);
int gx_scale(void *field, int n, int dtype, double a, void *stream);                   /* Field::operator*=(value) */
int gx_axpy(void *y, const void *x, int n, int dtype, double a, void *stream);         /* Field::addScaled :355-362 */
int gx_axpy_fourier(void *y, const void *x, int n, int dtype, double a, void *stream); /* MultiLevelField::addScaled :194-212 */
int gx_mul_mask(void *field, const unsigned char *mask_n3, int n, int dtype, void *stream); /* operator*=(Field<char>) :306-318 */
/* MultiLevelField::innerProduct for one level (:248-294): sum_modes w Re(conj(a) b), weights fourier.hpp:116-147 */
int gx_dot_fourier(const void *a, const void *b, int n, int dtype, double *host_out, void *stream);
/* Field::innerProduct in real space over the n^3 cells (field.hpp:365-377; the stale-padding quirk is dropped) */
int gx_dot_real(const void *a, const void *b, int n, int dtype, double *host_out, void *stream);
int gx_get_mode(const void *field, int n, int dtype, int kx, int ky, int kz, double *re, double *im); /* getFourierCoefficient */
int gx_set_mode(void *field, int n, int dtype, int kx, int ky, int kz, double re, double im);         /* setFourierCoefficient */

/* ---- zoom levels: fields on grids of different resolution, windows, flagged-cell masks ----------------------------- */
typedef struct {
  int n;                  /* Grid::size */
  double cell_size;       /* Grid::cellSize */
  double offset_lower[3]; /* Grid::offsetLower */
  double domain_size;     /* Grid::periodicDomainSize */
} gx_grid_desc;           /* the const metadata of grids::Grid<T>, grid.hpp:84-93 */
/* Field::addFieldFromDifferentGrid (field.hpp:804-812; evaluator.hpp:31-57), both fields in real space:
   coarser source: every dst cell the source covers += tricubic interpolation of the source at the cell centre
                   (evaluateInterpolated field.hpp:483-508, makeTricubicInterpolator :598-617 -- wrapped stencil if the source
                   spans the box, clamped otherwise; LocalUnitTricubicApproximation tricubic.hpp:54-77);
   finer source  : every dst cell whose first or last sub-cell the source holds += mean of the sub-cells it holds
                   (SubSampleEvaluator evaluator.hpp:168-205, SubSampleGrid::forEachSubcell virtualgrid.hpp:570-600). */
int gx_interp_add(void *dst, const gx_grid_desc *dst_grid, const void *src, const gx_grid_desc *src_grid, int dtype,
                  void *stream);
/* The deinterpolation loop of generateMultilevelCovectorFromHiresCovector (multilevelgrid.hpp:283-297):
   dst_coarse += sum over fine cells of Field::deInterpolate(centroid, value * scale) (field.hpp:420-480), i.e. the transpose
   of the tricubic interpolation (tricubic.hpp:450-549), gathered per coarse cell (deterministic, no atomics). */
int gx_deinterp_add(void *dst_coarse, const gx_grid_desc *coarse_grid, const void *src_fine, const gx_grid_desc *fine_grid,
                    int dtype, double scale, void *stream);
/* Field::setZeroInsideWindow / setZeroOutsideWindow (field.hpp:774-795) for window = window_grid->getWindow()
   (grid.hpp:502-504): zero the real-space cells whose centroid is inside (inside != 0) / outside the window. */
int gx_zero_window(void *field, const gx_grid_desc *grid, int dtype, const gx_grid_desc *window_grid, int inside,
                   void *stream);
/* flagged cells as a device mask of n^3 bytes (1 = flagged): sum over flagged cells, and
   x <- flagged ? (x - shift) * scale : 0 (windowOperator / varianceOperator, quadraticmodification.hpp:172-221) */
int gx_masked_sum(const void *field, const unsigned char *mask_n3, int n, int dtype, double *host_out, void *stream);
int gx_masked_shift_scale(void *field, const unsigned char *mask_n3, int n, int dtype, double shift, double scale,
                          void *stream);
/* potential_form == 0: F <- scale * F / k^2 on the independent modes, 0 at k = 0 (VelocityModification::computeVelocity,
   real-space-derivative branch, linearmodification.hpp:242-262; AngMomentumModification :386-411).
   potential_form != 0: cosmology::densityToPotential (parameters.hpp:208-225) as written: F <- scale * F * (1/k^2) with the
   reciprocal forced to 0 where the ky component is 0 (the reference compares its lambda argument `k2` = ky, not k^2). */
int gx_scale_inv_k2(void *field, int n, int dtype, double grid_length, double scale, int potential_form, void *stream);

/* ---- Zeldovich (ZeldovichParticleGenerator::calculateOffsetFields, zeldovich.hpp:152-181,248-251) ------- */
/* psi_axis(k) = i k_axis delta_k / k^2 (0 at k=0 and where k_axis is the Nyquist mode), independent slots
   only, into a separate field (generateNewFourierFields, fourier.hpp:253-274). axis: 0,1,2. */
int gx_zeldovich_fourier(const void *delta_k, void *psi_k, int n, int dtype, double grid_length, int axis,
                         void *stream);
/* The same followed by toReal() (mirror + c2r + 1/n^1.5) with the operator, the mirror and the scaling fused
   into the FFT passes: psi = real-space displacement component `axis`. delta_k is not modified. */
int gx_zeldovich_offset(const void *delta_k, void *psi, int n, int dtype, double grid_length, int axis,
                        void *stream);
/* ZeldovichParticleEvaluator::getParticle for a unigrid (zeldovich.hpp:87-119, generator.hpp:49-53,
   grid.hpp:519-531,656-669): for cell id = (x*n+y)*n+z,
     pos[axis] = wrap(centroid_axis + psi), vel[axis] = psi * vel_ratio.
   Output arrays are strided so that one call fills one component of an interleaved particle array:
   out[(id)*stride_elems]. out_dtype GX_F32/GX_F64. `id0`/`count` select a range of cells. */
int gx_emit_component(const void *psi, int n, int dtype, int axis, double cell_size, double offset_lower,
                      double domain_size, double vel_ratio, void *out_pos, void *out_vel, int out_dtype,
                      long stride_elems, long long id0, long long count, void *stream);

/* The device-side gather behind the particle mappers (TwoLevelParticleMapper::dereferenceIterator ->
   ParticleEvaluator::getParticle, twolevelmapper.hpp:846-852, generator.hpp:49-53): the evaluation of gx_emit_component at
   a DEVICE list of `count` cell indices of the grid (in mapper order); out[i*stride_elems] receives particle i. */
int gx_emit_cells(const void *psi, int n, int dtype, int axis, double cell_size, double offset_lower,
                  double domain_size, double vel_ratio, const long long *cells, long long count, void *out_pos,
                  void *out_vel, int out_dtype, long stride_elems, void *stream);

/* ---- whole unigrid path (ICGenerator::done -> write, ic.hpp:1633-1648,1275-1331) ------------------------------ */
typedef struct {
  int n;                   /* base_grid n */
  int dtype;               /* GX_F64, or GX_F32 for the reference's -f mode */
  int out_dtype;           /* particle output type (the reference's writers store float32) */
  int noise_mode;          /* GX_NOISE_* */
  int reverse_order;       /* seedfourier_reverse */
  int fix_power;           /* fix_power */
  unsigned long seed;
  double grid_length;      /* Grid::thisGridSize */
  double cell_size;        /* Grid::cellSize */
  double domain_size;      /* Grid::periodicDomainSize */
  double offset_lower[3];  /* Grid::offsetLower */
  double vel_ratio;        /* zeldovichVelocityToOffsetRatio, parameters.hpp:83-106 */
} gx_ic_params;
/* delta_field, psi_field: two device fields of gx_field_bytes(n,dtype). stage: device buffer of stage_bytes used as
   two halves, each holding pos+vel of a whole number of x-planes of one component.
   host_pos/host_vel (may be NULL: results stay in the staging buffer, nothing is copied): component-major arrays
   out[axis][id] of out_dtype; if host_capacity_particles < n^3 they are used as a ring of 3*capacity elements.
   Synchronous with respect to the device->host copies; kernels run on `stream`. Both fields are scratch on return: the
   schedule forms delta_k / k^2 once (the factor sqrt(C(k)) / k^2 is applied while the noise is written), emits psi_y, then
   shares one x-transform of delta_k / k^2 between psi_y and psi_z (5 column passes per IC instead of 6) and lets the last
   FFT pass write particles instead of real-space fields. */
int gx_unigrid_ic(const gx_ic_params *p, const gx_spectrum *spectrum, void *delta_field, void *psi_field, void *stage,
                  size_t stage_bytes, void *host_pos, void *host_vel, size_t host_capacity_particles, void *stream);

/* ---- slab decomposition over G GPUs (one process per GPU) ----------------------------------------------
   The reference is one process: FieldFourierManager::performTransform (fourier.hpp:479-518) transforms the whole field.
   Here rank r of `world` (a power of two dividing n, P = n/world) holds
     Fourier space: the ky-slab [kx:n][ky_local:P][kz:n/2+1], ky = r*P + ky_local
     real space   : the x-slab  [x_local:P][y:n][z:n+2],       x  = r*P + x_local
   and a 3-D transform is local x-pass | all-to-all | local y-pass | local z-pass (reverse order for r2c). The exchange
   buffer is [peer][x_local][ky_local][kz] (each peer block contiguous), which the x-pass produces/consumes in natural
   order and the y-pass addresses directly, so there is no pack/unpack sweep. The all-to-all and the all-gather of the
   two mirror planes are issued by the host layer (NCCL). Every function works on this rank's slab only. */
size_t gx_slab_bytes(int n, int dtype, int world);
/* RandomFieldGenerator::draw for this rank's slab: Fourier orders fill the ky-slab; GX_NOISE_REAL fills the x-slab in
   real space (follow with the slab r2c and gx_slab_zero_dc on rank 0). Bit-identical to gx_noise_draw slot for slot. */
int gx_slab_noise_draw(void *slab, int n, int dtype, int world, int rank, unsigned long seed, int mode,
                       int reverse_order, unsigned long long *stream_offset_io, void *stream);
/* The same for GX_NOISE_FOURIER_PARALLEL in phases, so that the chunk parse (pass A of the generator: the per-chunk transfer
   functions the scan needs, ~40 % of the noise time) is computed once across the ranks instead of once per rank:
     gx_slab_noise_pass_a  : this rank's share of the chunks -> fn_local (device, gx_slab_noise_fn_bytes(n, world) bytes)
     (host layer)          : all-gather of fn_local into fn_all (world * that many bytes, rank-major)
     gx_slab_noise_finish  : scan + delivery of the chunks that can reach this rank's slab. Bit-identical to gx_slab_noise_draw. */
size_t gx_slab_noise_fn_bytes(int n, int world);
int gx_slab_noise_pass_a(void *fn_local, int n, int dtype, int world, int rank, unsigned long seed, void *stream);
int gx_slab_noise_finish(void *slab, const void *fn_all, int n, int dtype, int world, int rank, unsigned long seed,
                         int reverse_order, void *stream);
/* The same draw with every chunk generated on exactly ONE rank (shells split by generator work) and every mode stored
   straight into its owner's slab through peer-mapped pointers (peer_slabs[r], from gx_ipc_open; peer_slabs[rank] == slab):
   the delivery kernel is the routing step, over NVLink. spectrum != NULL folds Field::applyTransferFunction(0.5) and the
   division by k^2 into the stores (as gx_unigrid_ic does). scratch: idle device memory for the staged draws (>= one
   shell's worth; NULL: allocated for the call). Bracket with barriers across the ranks. Bit-identical to gx_noise_draw. */
int gx_slab_noise_staged(void *slab, int n, int dtype, int world, int rank, unsigned long seed, int reverse_order,
                         void *const *peer_slabs, const gx_spectrum *spectrum, double grid_length, void *scratch,
                         size_t scratch_bytes, void *stream);
int gx_slab_zero_dc(void *slab, int dtype, void *stream);
/* Verification hook (no reference counterpart): device_accum9 != NULL makes every particle emit that follows
   (gx_unigrid_ic, gx_slab_z_emit) add sum(pos), sum(vel), sum(vel^2) of what it wrote to device_accum9[3*axis + 0..2]
   (9 doubles on the device, zeroed by the caller); NULL switches it off. bench.py prints the sums of one extra IC outside
   the timed loop, so an N-rank run can be compared with the 1-rank run of the same workload. */
int gx_emit_checksum(double *device_accum9);
int gx_slab_unit_variance(void *slab, int n, int dtype, int world, int rank, void *stream);
/* Field::applyTransferFunction on the ky-slab; divide_k2 != 0 leaves delta_k / k^2 (see gx_unigrid_ic) */
int gx_slab_apply_transfer(void *slab, int n, int dtype, int world, int rank, double grid_length, const gx_spectrum *s,
                           double power, int divide_k2, void *stream);
/* planes_local[n][P][2] <- the kz = 0 and kz = n/2 planes of the ky-slab (input of the mirror all-gather) */
int gx_slab_extract_planes(const void *slab, void *planes_local, int n, int dtype, int world, void *stream);
/* inverse x-pass of the ky-slab with ensureFourierModesAreMirrored (fourier.hpp:331-357) resolved on load from the
   all-gathered planes [peer][n][P][2]; op 0: mirror only (toReal of any field), 1: times i k_axis (slab holds
   delta_k/k^2), 2: times i k_axis / k^2 (zeldovich.hpp:170-176). out may not alias slab. */
int gx_slab_fft_x_mirror(const void *slab, const void *planes_all, void *out, int n, int dtype, int world, int rank,
                         double grid_length, int op, int axis, void *stream);
/* gx_slab_fft_x_mirror with the FFT transpose fused into the pass: every output element is stored straight into the
   exchange buffer of the rank that owns its x plane (peer_exchange[r]: peer-mapped buffers of gx_slab_bytes each, laid out
   [source rank][x_local][ky_local][kz]; peer_exchange[rank] is this rank's own) -- no send buffer, no all-to-all.
   Bracket with barriers across the ranks; follow with gx_slab_fft_y / gx_slab_fft_y_op reading the exchange layout. */
int gx_slab_fft_x_mirror_push(const void *slab, const void *planes_all, void *const *peer_exchange, int n, int dtype,
                              int world, int rank, double grid_length, int op, int axis, void *stream);
/* inverse y-pass with i k_axis (axis 1 or 2) applied on load: psi_y and psi_z from ONE x-transformed, exchanged
   G = delta_k / k^2 (zeldovich.hpp:152-181; neither factor depends on kx) */
int gx_slab_fft_y_op(const void *src, void *dst, int n, int dtype, int world, int src_exchange_layout, double grid_length,
                     int axis, void *stream);
int gx_slab_fft_x(void *slab, int n, int dtype, int world, int dir, int scale, void *stream); /* plain x-pass, in place; scale: / n^1.5 */
/* y-pass over the P local x-planes; either side may be the exchange layout instead of [x_local][y][kz] */
int gx_slab_fft_y(const void *src, void *dst, int n, int dtype, int world, int dir, int src_exchange_layout,
                  int dst_exchange_layout, void *stream);
int gx_slab_fft_z_r2c(void *slab, int n, int dtype, int world, void *stream);  /* rows, real -> half-complex, unscaled */
int gx_slab_fft_z_c2r(void *slab, int n, int dtype, int world, void *stream);  /* rows, half-complex -> real, / n^1.5 */
/* last inverse pass fused with ZeldovichParticleEvaluator::getParticle for local x-planes [plane0, plane0+nplanes):
   element i of out_pos/out_vel is particle id (rank*P + plane0)*n*n + i */
int gx_slab_z_emit(const void *slab, int n, int dtype, int world, int rank, int axis, double cell_size,
                   double offset_lower, double domain_size, double vel_ratio, int plane0, int nplanes, void *out_pos,
                   void *out_vel, int out_dtype, void *stream);

/* ---- per-kernel-class device timing for the roofline report (off by default) ---------------------------------- */
int gx_timing_enable(int on);
int gx_timing_report(char *buf, int buflen); /* lines "name count total_ms" */

#ifdef __cplusplus
}
#endif
#endif

// Zeldovich displacement with everything fused into the three inverse-FFT passes.
//
// Reference: ZeldovichParticleGenerator::calculateOffsetFields (zeldovich.hpp:152-181,248-251) builds three new
// Fourier fields through generateNewFourierFields (fourier.hpp:253-274), then each toReal() runs
// ensureFourierModesAreMirrored (:331-357), the c2r plan, unpadAfterFFTWRealTransform (:389-405) and the
// divide-by-N^1.5 sweep (:512-513); ZeldovichParticleEvaluator::getParticle (zeldovich.hpp:87-119) later reads the
// three real fields cell by cell. Here, per component:
//   x-pass : reads delta_k ONCE, applies i k_j / k^2 (Nyquist/DC zeroing) and the mirror on load, writes the
//            half-transformed field                                                                   (2 B)
//   y-pass : in place                                                                              (2 B)
//   z-pass : c2r + 1/N^1.5 and either the real field or directly pos = wrap(centroid + psi), vel = psi * ratio
//            in the output type                                                                     (<= 2 B)
// The arithmetic of each element is the unfused path's (spectral.cu ZeldovichF, mirror_kernel, particles.cu), so both
// paths give bit-identical results.
#include "../../include/genetic_b200.h"
#include "spectral.cuh"
#include "fft.cuh"
#include "zeldovich.cuh"
#include "zeldovich_ops.cuh"

namespace gx {

template <typename T, bool PREDIV> struct OpZeldovichMirror {
  const cplx<T> *delta;
  int n, nc, axis;
  T kmin;
  __device__ __forceinline__ cplx<T> operator()(cplx<T> d, int ix, int iy, int iz) const {
    const int h = n >> 1;
    bool conj = false;
    if ((iz == 0 || iz == h) && !slot_is_independent(n, ix, iy, iz)) {
      // ensureFourierModesAreMirrored: this slot is the conjugate of (-kx, -ky) on the same kz plane
      ix = (n - ix) & (n - 1);
      iy = (n - iy) & (n - 1);
      d = delta[((size_t) ix * n + iy) * nc + iz];
      conj = true;
    }
    const int ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
    cplx<T> r = mk<T>(0, 0);
    if (PREDIV) {
      // delta_k / k^2 was formed by the transfer sweep: only k_axis is needed here, and k^2 != 0 is an integer test
      // (k = integer * kMin), so the three products and two sums of slot_k drop out of every element of the pass
      if (ia != h && (ix | iy | iz) != 0) {
        const T ka = mul_rn((T) (axis == 2 ? iz : signed_mode(ia, n)), kmin);
        r.x = mul_rn(-d.y, ka); r.y = mul_rn(d.x, ka);
      }
    } else {
      T kx, ky, kz, k2;
      slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
      const T ka = axis == 0 ? kx : (axis == 1 ? ky : kz);
      if (ia != h && k2 != (T) 0) { r.x = mul_rn(div_rn(-d.y, k2), ka); r.y = mul_rn(div_rn(d.x, k2), ka); }
    }
    // the mirror leaves conj(value) in dependent slots and a negated imaginary part in self-conjugate ones
    const bool self = (iz == 0 || iz == h) && (ix == 0 || ix == h) && (iy == 0 || iy == h);
    if (conj || self) r.y = -r.y;
    return r;
  }
};

// ---- pieces of the shared-transform schedule (pipeline.cu): the x-transform of G = delta_k / k^2 is computed once, in
// place, and serves both psi_y (i k_y does not depend on kx: applied on the load side of its y-pass) and psi_z (i k_z on
// the store side of its y-pass), so an IC takes 5 column passes instead of 6 -- and every one of them is of the cheap kind
// (x-passes in place, the out-of-place ones along y). The Hermitian mirror commutes with the diagonal operators on the
// dependent slots (mirror(i k G) = i k mirror(G)), and the imaginary parts it leaves on self-conjugate entries are dropped
// by the c2r pass either way, so G's two planes are mirrored once, explicitly, before the shared x-pass.
template <typename O> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<double> { using type = double2; };

template <typename T, typename O> struct StoreEmit {
  O *pos, *vel;
  long long id0;
  int n, axis;
  T scale, dx, lower, half_dx, domain, ratio;
  using P2 = typename Pair<O>::type;
  struct Ctx { P2 *pos, *vel; T cen; };
  // getCentroidFromCoordinate: coord*cellSize; += offsetLower; += cellSize/2 (three separately rounded operations)
  __device__ __forceinline__ T centre(int c) const { return add_rn(add_rn(mul_rn((T) c, dx), lower), half_dx); }
  __device__ __forceinline__ Ctx begin(long row) const {
    Ctx c;
    const long long first = (long long) row * n - id0;  // index of particle (row, z = 0) in the output arrays (even)
    c.pos = pos ? reinterpret_cast<P2 *>(pos + first) : nullptr;
    c.vel = vel ? reinterpret_cast<P2 *>(vel + first) : nullptr;
    const int x = (int) (row / n);
    c.cen = centre(axis == 0 ? x : (int) (row - (long) x * n));  // unused for axis 2
    return c;
  }
  __device__ __forceinline__ void put(const Ctx &c, int k, cplx<T> z) const {
    const T p0 = z.x * scale, p1 = z.y * scale;
    if (c.pos) {
      const T c0 = axis == 2 ? centre(2 * k) : c.cen, c1 = axis == 2 ? centre(2 * k + 1) : c.cen;
      P2 o;
      o.x = (O) wrap_coordinate<T>(add_rn(p0, c0), domain);
      o.y = (O) wrap_coordinate<T>(add_rn(p1, c1), domain);
      c.pos[k] = o;
    }
    if (c.vel) {
      P2 o; o.x = (O) mul_rn(p0, ratio); o.y = (O) mul_rn(p1, ratio);
      c.vel[k] = o;
    }
  }
};

// order 0: x-pass (fused load, out of place) then y-pass in place.
// order 1: y-pass (fused load, out of place) then x-pass in place. The two 1-D transforms commute; which one carries the
// out-of-place operator matters for HBM: a 2048^3 x-line visits 2048 distinct 2 MB pages, and reading delta_k while
// writing the work field doubles that footprint, whereas a y-pass tile spans 16 pages.
template <typename T, int N, bool PREDIV>
static int zeldovich_xy_n(const void *delta_k, void *work, double L, int axis, int order, cudaStream_t st) {
  using namespace fft;
  using C = cplx<T>;
  const int nc = N / 2 + 1;
  OpZeldovichMirror<T, PREDIV> op;
  op.delta = (const C *) delta_k; op.n = N; op.nc = nc; op.axis = axis;
  op.kmin = (T) (2. * M_PI / (double) (T) L);
  if (order == 1) {
    int rc = launch_cols<T, N, 1, PassY>((const C *) delta_k, (C *) work, N, N, nc, op, OpIdentity<T>(), st);
    if (rc) return rc;
    return launch_cols<T, N, 1, PassX>((const C *) work, (C *) work, N, N, nc, OpIdentity<T>(), OpIdentity<T>(), st);
  }
  int rc = launch_cols<T, N, 1, PassX>((const C *) delta_k, (C *) work, N, N, nc, op, OpIdentity<T>(), st);
  if (rc) return rc;
  return launch_cols<T, N, 1, PassY>((const C *) work, (C *) work, N, N, nc, OpIdentity<T>(), OpIdentity<T>(), st);
}

template <typename T, int N> static int z_real_n(void *work, long row0, long nrows, cudaStream_t st) {
  fft::StoreRealScaled<T> s;
  s.data = (T *) work; s.scale = (T) (1.0 / pow((double) N, 1.5)); s.n = N;
  return fft::launch_rows_c2r<T, N>((const T *) work, row0, nrows, s, st);
}

template <typename T, typename O, int N>
static int z_emit_n(const void *work, int axis, const EmitGeo &g, long row0, long nrows, void *pos, void *vel, cudaStream_t st) {
  StoreEmit<T, O> s;
  s.pos = (O *) pos; s.vel = (O *) vel; s.id0 = (long long) row0 * N; s.n = N; s.axis = axis;
  s.scale = (T) (1.0 / pow((double) N, 1.5));
  s.dx = (T) g.cell_size; s.lower = (T) g.offset_lower; s.half_dx = (T) g.cell_size / (T) 2;
  s.domain = (T) g.domain_size; s.ratio = (T) g.vel_ratio;
  return fft::launch_rows_c2r<T, N>((const T *) work, row0, nrows, s, st);
}

// G <- x-pass(G) in place (after the explicit mirror); then, from the x-transformed G:
//   axis 1: W <- y-pass(i k_y G) out of place;  axis 2: G <- y-pass(G) in place, times i k_z as it is stored
template <typename T, int N> static int shared_x_n(void *g, cudaStream_t st) {
  using namespace fft;
  using C = cplx<T>;
  return launch_cols<T, N, 1, PassX>((const C *) g, (C *) g, N, N, N / 2 + 1, OpIdentity<T>(), OpIdentity<T>(), st);
}
template <typename T, int N> static int shared_y_n(void *g, void *w, double L, int axis, cudaStream_t st) {
  using namespace fft;
  using C = cplx<T>;
  const T kmin = (T) (2. * M_PI / (double) (T) L);
  if (axis == 1) {
    OpIKy<T> op;
    op.n = N; op.kmin = kmin;
    return launch_cols<T, N, 1, PassY>((const C *) g, (C *) w, N, N, N / 2 + 1, op, OpIdentity<T>(), st);
  }
  OpIKz<T> op;
  op.n = N; op.kmin = kmin;
  return launch_cols<T, N, 1, PassY>((const C *) g, (C *) g, N, N, N / 2 + 1, OpIdentity<T>(), op, st);
}
template <typename T> static int shared_x_t(void *g, int n, cudaStream_t st) { GX_N_SWITCH(n, (shared_x_n<T, N>(g, st))); }
template <typename T> static int shared_y_t(void *g, void *w, int n, double L, int axis, cudaStream_t st) {
  GX_N_SWITCH(n, (shared_y_n<T, N>(g, w, L, axis, st)));
}

template <typename T, bool PREDIV>
static int zeldovich_xy_t(const void *d, void *w, int n, double L, int axis, int order, cudaStream_t st) {
  GX_N_SWITCH(n, (zeldovich_xy_n<T, N, PREDIV>(d, w, L, axis, order, st)));
}
template <typename T> static int z_real_t(void *w, int n, long row0, long nrows, cudaStream_t st) {
  GX_N_SWITCH(n, (z_real_n<T, N>(w, row0, nrows, st)));
}
template <typename T, typename O>
static int z_emit_t(const void *w, int n, int axis, const EmitGeo &g, long row0, long nrows, void *pos, void *vel,
                    cudaStream_t st) {
  GX_N_SWITCH(n, (z_emit_n<T, O, N>(w, axis, g, row0, nrows, pos, vel, st)));
}

int zeldovich_xy(const void *delta_k, void *work, int n, int dtype, double grid_length, int axis, int prediv,
                 cudaStream_t st) {
  if (axis < 0 || axis > 2) return fail("axis must be 0, 1 or 2");
  if (delta_k == work) return fail("zeldovich: the work field must not alias delta_k");
  if (int rc = fft::ensure_twiddles()) return rc;
  static int order = -1;
  if (order < 0) {
    const char *e = getenv("GX_ZELDOVICH_ORDER");  // development switch: "xy" or "yx"
    order = (e && e[0] == 'x') ? 0 : 1;
  }
  if (dtype == GX_F64)
    return prediv ? zeldovich_xy_t<double, true>(delta_k, work, n, grid_length, axis, order, st)
                  : zeldovich_xy_t<double, false>(delta_k, work, n, grid_length, axis, order, st);
  if (dtype == GX_F32)
    return prediv ? zeldovich_xy_t<float, true>(delta_k, work, n, grid_length, axis, order, st)
                  : zeldovich_xy_t<float, false>(delta_k, work, n, grid_length, axis, order, st);
  return fail("bad dtype");
}

int zeldovich_z_real(void *work, int n, int dtype, long row0, long nrows, cudaStream_t st) {
  if (dtype == GX_F64) return z_real_t<double>(work, n, row0, nrows, st);
  if (dtype == GX_F32) return z_real_t<float>(work, n, row0, nrows, st);
  return fail("bad dtype");
}

// ---- verification hook: sums over everything the emit writes ----------------------------------------------------------
// bench.py prints them (one extra IC outside the timed loop) so that the particles of an N-rank run can be compared with the
// 1-rank run of the same workload: per axis sum(pos), sum(vel), sum(vel^2), accumulated in double.
static double *g_emit_checksum = nullptr;  // device, 9 doubles: [axis][pos_sum, vel_sum, vel_sumsq]

template <typename O> __global__ void emit_checksum_kernel(const O *pos, const O *vel, long long count, double *acc) {
  double sp = 0, sv = 0, sq = 0;
  for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < count; i += (long long) gridDim.x * blockDim.x) {
    if (pos) sp += (double) pos[i];
    if (vel) { const double v = (double) vel[i]; sv += v; sq += v * v; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_down_sync(0xffffffffu, sp, o);
    sv += __shfl_down_sync(0xffffffffu, sv, o);
    sq += __shfl_down_sync(0xffffffffu, sq, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(acc, sp); atomicAdd(acc + 1, sv); atomicAdd(acc + 2, sq); }
}

int zeldovich_z_emit(const void *work, int n, int dtype, int axis, const EmitGeo &geo, long row0, long nrows,
                     void *out_pos, void *out_vel, int out_dtype, cudaStream_t st) {
  int rc;
  if (dtype == GX_F64 && out_dtype == GX_F32) rc = z_emit_t<double, float>(work, n, axis, geo, row0, nrows, out_pos, out_vel, st);
  else if (dtype == GX_F64 && out_dtype == GX_F64) rc = z_emit_t<double, double>(work, n, axis, geo, row0, nrows, out_pos, out_vel, st);
  else if (dtype == GX_F32 && out_dtype == GX_F32) rc = z_emit_t<float, float>(work, n, axis, geo, row0, nrows, out_pos, out_vel, st);
  else if (dtype == GX_F32 && out_dtype == GX_F64) rc = z_emit_t<float, double>(work, n, axis, geo, row0, nrows, out_pos, out_vel, st);
  else return fail("bad dtype");
  if (rc || !g_emit_checksum) return rc;
  const long long count = (long long) nrows * n;
  if (out_dtype == GX_F32)
    emit_checksum_kernel<float><<<148 * 8, 256, 0, st>>>((const float *) out_pos, (const float *) out_vel, count, g_emit_checksum + 3 * axis);
  else
    emit_checksum_kernel<double><<<148 * 8, 256, 0, st>>>((const double *) out_pos, (const double *) out_vel, count, g_emit_checksum + 3 * axis);
  GX_LAUNCH_CHECK();
  return 0;
}

int zeldovich_shared_x(void *g, int n, int dtype, cudaStream_t st) {
  if (int rc = fft::ensure_twiddles()) return rc;
  int rc = gx_mirror(g, n, dtype, st);  // ensureFourierModesAreMirrored on G's kz = 0 and Nyquist planes, in place
  if (rc) return rc;
  if (dtype == GX_F64) return shared_x_t<double>(g, n, st);
  if (dtype == GX_F32) return shared_x_t<float>(g, n, st);
  return fail("bad dtype");
}

int zeldovich_shared_y(void *g, void *work, int n, int dtype, double grid_length, int axis, cudaStream_t st) {
  if (axis != 1 && axis != 2) return fail("zeldovich_shared_y: axis must be 1 or 2");
  if (g == work) return fail("zeldovich: the work field must not alias delta_k");
  if (dtype == GX_F64) return shared_y_t<double>(g, work, n, grid_length, axis, st);
  if (dtype == GX_F32) return shared_y_t<float>(g, work, n, grid_length, axis, st);
  return fail("bad dtype");
}

}  // namespace gx

extern "C" int gx_zeldovich_offset(const void *delta_k, void *psi, int n, int dtype, double grid_length, int axis,
                                   void *stream) {
  cudaStream_t st = (cudaStream_t) stream;
  if (!gx::is_pow2(n)) {  // mixed-radix sizes: the operator as a sweep, then the plain inverse transform
    if (int rc0 = gx_zeldovich_fourier(delta_k, psi, n, dtype, grid_length, axis, stream)) return rc0;
    return gx_fft_inverse(psi, n, dtype, stream);
  }
  int rc = gx::zeldovich_xy(delta_k, psi, n, dtype, grid_length, axis, 0, st);
  if (rc) return rc;
  return gx::zeldovich_z_real(psi, n, dtype, 0, (long) n * n, st);
}

// device_accum9 != NULL: every particle emit that follows (gx_unigrid_ic, gx_slab_z_emit) also adds sum(pos), sum(vel),
// sum(vel^2) of what it wrote to device_accum9[3 * axis + 0..2] (zero it first); NULL switches the hook off.
extern "C" int gx_emit_checksum(double *device_accum9) {
  gx::g_emit_checksum = device_accum9;
  return 0;
}

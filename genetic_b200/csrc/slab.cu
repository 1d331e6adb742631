// Slab-decomposed passes: the per-GPU half of the multi-GPU 3-D FFT and Zeldovich path.
//
// The reference is a single process (no MPI/NCCL); its FieldFourierManager::performTransform (fourier.hpp:479-518)
// transforms the whole field at once. Here G ranks (one per GPU, G a power of two dividing n, P = n / G) hold
//   Fourier space : the ky-slab  F_r[kx : n][ky_local : P][kz : n/2+1],  ky = r P + ky_local   (whole kx lines)
//   real space    : the x-slab   R_r[x_local : P][y : n][z : n+2],       x  = r P + x_local    (particle ids are
//                   contiguous per rank, as OneLevelParticleMapper numbers them, onelevelmapper.hpp:38-41)
// so every 3-D transform is: local x-pass | ONE all-to-all | local y-pass | local z-pass (inverse; reversed for the
// forward transform). The exchange buffer is laid out [peer][x_local][ky_local][kz]: the x-pass output (inverse) is
// already in that order for sending -- the block for peer s is the contiguous range x in [sP, (s+1)P) -- and the y-pass
// reads the received blocks in place through a blocked line map (fft.cuh LineMap), so no pack or unpack sweep exists.
// The all-to-all itself is issued by the host layer (genetic_b200/slab.py, torch.distributed / NCCL over NVLink).
//
// The Hermitian mirror (ensureFourierModesAreMirrored, fourier.hpp:331-357) couples ky with -ky, i.e. two different
// ranks, but only on the planes kz = 0 and kz = n/2: the ranks all-gather those two planes (2 n^2 complex) once and the
// x-pass resolves dependent slots from the gathered copy on load.
#include "../../include/genetic_b200.h"
#include "spectral.cuh"
#include "fft.cuh"
#include "zeldovich.cuh"
#include "zeldovich_ops.cuh"
#include <algorithm>

namespace gx {

static int check_slab(int n, int world, int rank) {
  if (!is_pow2(n) || n < 4 || n > 4096) return fail("grid size must be a power of two in [4,4096]");
  if (world < 1 || !is_pow2(world) || n % world != 0) return fail("slab: world must be a power of two dividing n");
  if (rank < 0 || rank >= world) return fail("slab: bad rank");
  return 0;
}

// MODE 0: mirror only (general toReal). MODE 1: i k_axis * value (delta_k / k^2 precomputed). MODE 2: i k_axis / k^2.
template <typename T, int MODE> struct OpSlabMirror {
  const cplx<T> *planes;  // all-gathered [peer][n][P][2]: kz = 0 and kz = n/2 of every slab
  int n, axis, iy0, P, pshift;
  T kmin;
  __device__ __forceinline__ cplx<T> operator()(cplx<T> d, int ix, int iyl, int iz) const {
    const int h = n >> 1;
    int iy = iy0 + iyl;
    bool conj = false;
    if ((iz == 0 || iz == h) && !slot_is_independent(n, ix, iy, iz)) {
      ix = (n - ix) & (n - 1);
      iy = (n - iy) & (n - 1);
      const int s = iy >> pshift, j = iy & (P - 1);
      d = planes[(((size_t) s * n + ix) * P + j) * 2 + (iz == h ? 1 : 0)];
      conj = true;
    }
    cplx<T> r = d;
    if (MODE != 0) {
      T kx, ky, kz, k2;
      slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
      const int ia = axis == 0 ? ix : (axis == 1 ? iy : iz);
      const T ka = axis == 0 ? kx : (axis == 1 ? ky : kz);
      r = mk<T>(0, 0);
      if (ia != h && k2 != (T) 0) {
        if (MODE == 1) { r.x = mul_rn(-d.y, ka); r.y = mul_rn(d.x, ka); }
        else { r.x = mul_rn(div_rn(-d.y, k2), ka); r.y = mul_rn(div_rn(d.x, k2), ka); }
      }
    }
    const bool self = (iz == 0 || iz == h) && (ix == 0 || ix == h) && (iy == 0 || iy == h);
    if (conj || self) r.y = -r.y;
    return r;
  }
};

template <typename T> __global__ void extract_planes_kernel(const cplx<T> *slab, cplx<T> *planes, long rows, int nc, int h) {
  for (long i = blockIdx.x * (long) blockDim.x + threadIdx.x; i < rows; i += (long) gridDim.x * blockDim.x) {
    planes[2 * i] = slab[i * nc];
    planes[2 * i + 1] = slab[i * nc + h];
  }
}

// ---- x-pass on a ky-slab [n][P][nc] ---------------------------------------------------------------------------------
template <typename T, int N, int DIR, class LoadOp, class StoreOp>
static int slab_x_n(const void *src, void *dst, int P, LoadOp lop, StoreOp sop, cudaStream_t st) {
  using namespace fft;
  return launch_cols<T, N, DIR, PassX>((const cplx<T> *) src, (cplx<T> *) dst, N, P, N / 2 + 1, lop, sop, st);
}

template <typename T, int MODE>
static int slab_x_mirror_t(const void *src, const void *planes, void *dst, int n, int world, int rank, double L, int axis,
                           cudaStream_t st) {
  OpSlabMirror<T, MODE> op;
  op.planes = (const cplx<T> *) planes; op.n = n; op.axis = axis; op.P = n / world; op.iy0 = rank * op.P;
  op.pshift = 0; while ((1 << op.pshift) < op.P) ++op.pshift;
  op.kmin = (T) (2. * M_PI / (double) (T) L);
  const int P = op.P;
  GX_N_SWITCH(n, (slab_x_n<T, N, 1>(src, dst, P, op, fft::OpIdentity<T>(), st)));
}

// the same pass with the exchange fused into its store: every element goes straight to its owner's exchange buffer
template <typename T, int MODE>
static int slab_x_push_t(const void *src, const void *planes, void *const *peer_bufs, int n, int world, int rank, double L,
                         int axis, cudaStream_t st) {
  OpSlabMirror<T, MODE> op;
  op.planes = (const cplx<T> *) planes; op.n = n; op.axis = axis; op.P = n / world; op.iy0 = rank * op.P;
  op.pshift = 0; while ((1 << op.pshift) < op.P) ++op.pshift;
  op.kmin = (T) (2. * M_PI / (double) (T) L);
  const int P = op.P;
  const long nc = n / 2 + 1;
  fft::OpPushX<T> push;
  for (int r = 0; r < 8; ++r) push.peer[r] = (cplx<T> *) (r < world ? peer_bufs[r] : nullptr);
  push.pshift = op.pshift; push.pmask = P - 1;
  push.self_off = (long) rank * P * P * nc;
  push.line_stride = (long) P * nc;
  struct Limit {  // GX_PUSH_CTAS_PER_SM > 0: a persistent grid of that many CTAs per SM (see cols_cluster_limit)
    explicit Limit(int v) { fft::cols_cluster_limit() = v; }
    ~Limit() { fft::cols_cluster_limit() = 0; }
  };
  static int per_sm = -1;
  if (per_sm < 0) { const char *e = getenv("GX_PUSH_CTAS_PER_SM"); per_sm = e ? atoi(e) : 0; }
  int clusters = 0;
  if (per_sm > 0) {
    const int cl = n > 256 ? std::min(8, n / 256) : 1;
    clusters = std::max(1, per_sm * 148 / cl);
  }
  Limit guard(clusters);
  GX_N_SWITCH(n, (slab_x_n<T, N, 1>(src, nullptr, P, op, push, st)));
}

template <typename T> static int slab_x_plain_t(void *data, int n, int world, int dir, int scale, cudaStream_t st) {
  const int P = n / world;
  const T s = scale ? (T) (1.0 / pow((double) n, 1.5)) : (T) 1;
  if (dir < 0) { GX_N_SWITCH(n, (slab_x_n<T, N, -1>(data, data, P, fft::OpIdentity<T>(), fft::OpScale<T>(s), st))); }
  GX_N_SWITCH(n, (slab_x_n<T, N, 1>(data, data, P, fft::OpIdentity<T>(), fft::OpScale<T>(s), st)));
}

// ---- y-pass on an x-slab: P batches of [n][nc]; one side may be the exchange layout [peer][x_local][ky_local][nc] ----
template <typename T, int N, int DIR>
static int slab_y_n(const void *src, void *dst, int P, int src_blocked, int dst_blocked, cudaStream_t st) {
  using namespace fft;
  using PL = ColsPlan<T, N>;
  const int nc = N / 2 + 1;
  ColGeo g;
  g.nc = nc; g.tiles_per_batch = 0; g.ncols = nc;
  const LineMap plain = plain_map(nc, (long) N * nc);
  const LineMap blocked = blocked_map(nc, (long) P * nc, P, (long) P * P * nc);
  g.src = src_blocked ? blocked : plain;
  g.dst = dst_blocked ? blocked : plain;
  return launch_cols_cfg<T, N, PL::CL, PL::TC, PL::MINB, PL::SEL, DIR, PassY>((const cplx<T> *) src, (cplx<T> *) dst, g, P,
                                                                             OpIdentity<T>(), OpIdentity<T>(), st);
}

// inverse y-pass out of the exchange layout with i k_axis applied as the elements are loaded (axis 1 or 2)
template <typename T, int N, class LoadOp>
static int slab_y_op_n(const void *src, void *dst, int P, int src_blocked, LoadOp lop, cudaStream_t st) {
  using namespace fft;
  using PL = ColsPlan<T, N>;
  const int nc = N / 2 + 1;
  ColGeo g;
  g.nc = nc; g.tiles_per_batch = 0; g.ncols = nc;
  const LineMap plain = plain_map(nc, (long) N * nc);
  const LineMap blocked = blocked_map(nc, (long) P * nc, P, (long) P * P * nc);
  g.src = src_blocked ? blocked : plain;
  g.dst = plain;
  return launch_cols_cfg<T, N, PL::CL, PL::TC, PL::MINB, PL::SEL, 1, PassY>((const cplx<T> *) src, (cplx<T> *) dst, g, P, lop,
                                                                           OpIdentity<T>(), st);
}
template <typename T>
static int slab_y_op_t(const void *src, void *dst, int n, int world, int sb, double L, int axis, cudaStream_t st) {
  const int P = n / world;
  const T kmin = (T) (2. * M_PI / (double) (T) L);
  if (axis == 1) {
    OpIKy<T> op; op.n = n; op.kmin = kmin;
    GX_N_SWITCH(n, (slab_y_op_n<T, N>(src, dst, P, sb, op, st)));
  }
  OpIKz<T> op; op.n = n; op.kmin = kmin;
  GX_N_SWITCH(n, (slab_y_op_n<T, N>(src, dst, P, sb, op, st)));
}

template <typename T>
static int slab_y_t(const void *src, void *dst, int n, int world, int dir, int sb, int db, cudaStream_t st) {
  const int P = n / world;
  if (dir < 0) { GX_N_SWITCH(n, (slab_y_n<T, N, -1>(src, dst, P, sb, db, st))); }
  GX_N_SWITCH(n, (slab_y_n<T, N, 1>(src, dst, P, sb, db, st)));
}

// ---- z-passes on an x-slab ----------------------------------------------------------------------------------------------
template <typename T> static int slab_z_r2c_t(void *data, int n, int world, cudaStream_t st) {
  const long rows = (long) n * (n / world);
  GX_N_SWITCH(n, (fft::launch_rows_r2c<T, N>((T *) data, rows, st)));
}

}  // namespace gx

using namespace gx;

extern "C" {

size_t gx_slab_bytes(int n, int dtype, int world) {
  if (world < 1) return 0;
  return gx_field_bytes(n, dtype) / (size_t) world;
}

int gx_slab_apply_transfer(void *slab, int n, int dtype, int world, int rank, double grid_length, const gx_spectrum *s,
                           double power, int divide_k2, void *stream) {
  if (check_slab(n, world, rank)) return 1;
  const int P = n / world;
  return apply_transfer_ex(slab, n, dtype, grid_length, s, power, divide_k2, (cudaStream_t) stream, rank * P, P);
}

int gx_slab_unit_variance(void *slab, int n, int dtype, int world, int rank, void *stream) {
  if (check_slab(n, world, rank)) return 1;
  const int P = n / world;
  return unit_variance_slab(slab, n, dtype, rank * P, P, (cudaStream_t) stream);
}

int gx_slab_extract_planes(const void *slab, void *planes_local, int n, int dtype, int world, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  const long rows = (long) n * (n / world);
  const int nc = n / 2 + 1;
  const int grid = (int) std::min<long>((rows + 255) / 256, 148 * 8);
  cudaStream_t st = (cudaStream_t) stream;
  if (dtype == GX_F64) extract_planes_kernel<double><<<grid, 256, 0, st>>>((const double2 *) slab, (double2 *) planes_local, rows, nc, n / 2);
  else if (dtype == GX_F32) extract_planes_kernel<float><<<grid, 256, 0, st>>>((const float2 *) slab, (float2 *) planes_local, rows, nc, n / 2);
  else return fail("bad dtype");
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

int gx_slab_fft_x_mirror(const void *slab, const void *planes_all, void *out, int n, int dtype, int world, int rank,
                         double grid_length, int op, int axis, void *stream) {
  if (check_slab(n, world, rank)) return 1;
  if (op < 0 || op > 2) return fail("gx_slab_fft_x_mirror: op must be 0 (mirror), 1 (i k_axis) or 2 (i k_axis / k^2)");
  if (op != 0 && (axis < 0 || axis > 2)) return fail("axis must be 0, 1 or 2");
  if (int rc = fft::ensure_twiddles()) return rc;
  cudaStream_t st = (cudaStream_t) stream;
#define GX_SLAB_X(T)                                                                                              \
  switch (op) {                                                                                                  \
    case 0: return slab_x_mirror_t<T, 0>(slab, planes_all, out, n, world, rank, grid_length, axis, st);           \
    case 1: return slab_x_mirror_t<T, 1>(slab, planes_all, out, n, world, rank, grid_length, axis, st);           \
    default: return slab_x_mirror_t<T, 2>(slab, planes_all, out, n, world, rank, grid_length, axis, st);          \
  }
  if (dtype == GX_F64) { GX_SLAB_X(double) }
  if (dtype == GX_F32) { GX_SLAB_X(float) }
#undef GX_SLAB_X
  return fail("bad dtype");
}

int gx_slab_fft_x_mirror_push(const void *slab, const void *planes_all, void *const *peer_exchange, int n, int dtype,
                              int world, int rank, double grid_length, int op, int axis, void *stream) {
  if (check_slab(n, world, rank)) return 1;
  if (world > 8) return fail("gx_slab_fft_x_mirror_push: at most 8 ranks");
  if (op < 0 || op > 2) return fail("gx_slab_fft_x_mirror_push: op must be 0 (mirror), 1 (i k_axis) or 2 (i k_axis / k^2)");
  if (op != 0 && (axis < 0 || axis > 2)) return fail("axis must be 0, 1 or 2");
  if (!peer_exchange) return fail("gx_slab_fft_x_mirror_push: null peer table");
  for (int r = 0; r < world; ++r) if (!peer_exchange[r]) return fail("gx_slab_fft_x_mirror_push: null peer pointer");
  if (int rc = fft::ensure_twiddles()) return rc;
  cudaStream_t st = (cudaStream_t) stream;
#define GX_SLAB_XP(T)                                                                                              \
  switch (op) {                                                                                                   \
    case 0: return slab_x_push_t<T, 0>(slab, planes_all, peer_exchange, n, world, rank, grid_length, axis, st);    \
    case 1: return slab_x_push_t<T, 1>(slab, planes_all, peer_exchange, n, world, rank, grid_length, axis, st);    \
    default: return slab_x_push_t<T, 2>(slab, planes_all, peer_exchange, n, world, rank, grid_length, axis, st);   \
  }
  if (dtype == GX_F64) { GX_SLAB_XP(double) }
  if (dtype == GX_F32) { GX_SLAB_XP(float) }
#undef GX_SLAB_XP
  return fail("bad dtype");
}

int gx_slab_fft_y_op(const void *src, void *dst, int n, int dtype, int world, int src_exchange_layout, double grid_length,
                     int axis, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  if (axis != 1 && axis != 2) return fail("gx_slab_fft_y_op: axis must be 1 (i k_y) or 2 (i k_z)");
  if (src == dst && src_exchange_layout) return fail("gx_slab_fft_y_op: in-place needs the plain layout on both sides");
  if (int rc = fft::ensure_twiddles()) return rc;
  if (dtype == GX_F64) return slab_y_op_t<double>(src, dst, n, world, src_exchange_layout, grid_length, axis, (cudaStream_t) stream);
  if (dtype == GX_F32) return slab_y_op_t<float>(src, dst, n, world, src_exchange_layout, grid_length, axis, (cudaStream_t) stream);
  return fail("bad dtype");
}

int gx_slab_fft_x(void *slab, int n, int dtype, int world, int dir, int scale, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  if (int rc = fft::ensure_twiddles()) return rc;
  if (dtype == GX_F64) return slab_x_plain_t<double>(slab, n, world, dir, scale, (cudaStream_t) stream);
  if (dtype == GX_F32) return slab_x_plain_t<float>(slab, n, world, dir, scale, (cudaStream_t) stream);
  return fail("bad dtype");
}

int gx_slab_fft_y(const void *src, void *dst, int n, int dtype, int world, int dir, int src_exchange_layout,
                  int dst_exchange_layout, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  if (src == dst && src_exchange_layout != dst_exchange_layout)
    return fail("gx_slab_fft_y: in-place needs the same layout on both sides");
  if (int rc = fft::ensure_twiddles()) return rc;
  if (dtype == GX_F64) return slab_y_t<double>(src, dst, n, world, dir, src_exchange_layout, dst_exchange_layout, (cudaStream_t) stream);
  if (dtype == GX_F32) return slab_y_t<float>(src, dst, n, world, dir, src_exchange_layout, dst_exchange_layout, (cudaStream_t) stream);
  return fail("bad dtype");
}

int gx_slab_fft_z_r2c(void *slab, int n, int dtype, int world, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  if (int rc = fft::ensure_twiddles()) return rc;
  if (dtype == GX_F64) return slab_z_r2c_t<double>(slab, n, world, (cudaStream_t) stream);
  if (dtype == GX_F32) return slab_z_r2c_t<float>(slab, n, world, (cudaStream_t) stream);
  return fail("bad dtype");
}

int gx_slab_fft_z_c2r(void *slab, int n, int dtype, int world, void *stream) {
  if (check_slab(n, world, 0)) return 1;
  if (int rc = fft::ensure_twiddles()) return rc;
  return zeldovich_z_real(slab, n, dtype, 0, (long) n * (n / world), (cudaStream_t) stream);
}

int gx_slab_z_emit(const void *slab, int n, int dtype, int world, int rank, int axis, double cell_size,
                   double offset_lower, double domain_size, double vel_ratio, int plane0, int nplanes, void *out_pos,
                   void *out_vel, int out_dtype, void *stream) {
  if (check_slab(n, world, rank)) return 1;
  const int P = n / world;
  if (plane0 < 0 || nplanes < 0 || plane0 + nplanes > P) return fail("gx_slab_z_emit: plane range outside the slab");
  if (axis < 0 || axis > 2) return fail("axis must be 0, 1 or 2");
  if (int rc = fft::ensure_twiddles()) return rc;
  EmitGeo g;
  g.cell_size = cell_size; g.offset_lower = offset_lower; g.domain_size = domain_size; g.vel_ratio = vel_ratio;
  // rows are addressed by their global index x*n + y (the emit derives the cell coordinates from it)
  const size_t esz = dtype == GX_F64 ? 8 : 4;
  const long x0 = (long) rank * P;
  const char *virt = (const char *) slab - (size_t) x0 * n * (size_t) (n + 2) * esz;
  return zeldovich_z_emit(virt, n, dtype, axis, g, (x0 + plane0) * n, (long) nplanes * n, out_pos, out_vel, out_dtype,
                          (cudaStream_t) stream);
}

int gx_slab_zero_dc(void *slab, int dtype, void *stream) {  // tools::set_zero(fieldData[0]), randomfieldgenerator.hpp:195 (rank 0)
  return gx_memset_zero(slab, dtype == GX_F64 ? 8 : 4, stream);
}

}  // extern "C"

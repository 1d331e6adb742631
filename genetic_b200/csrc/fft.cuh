// Pass kernels of the 3-D FFT (templates; instantiated from fft.cu and the fused Zeldovich path in pipeline.cu).
//
// Column passes (the two strided axes) -- fft_cols_kernel
//   A tile is TC adjacent columns (TC * sizeof(complex) = 256 contiguous bytes) x the whole line of N points.
//   profiles/r1_segbench2048.txt: with a 33 MB line stride, 32-byte segments reach 1.5 TB/s, 64 B 3.0 TB/s,
//   128 B 5.7 TB/s and 256 B 6.5 TB/s, but only when several 64 KB tiles are resident per SM. A 2048-point f64
//   line of 16 columns is 512 KB, so the line is split over a thread-block CLUSTER of CL CTAs, decimation in time:
//   CTA q loads the rows j = q + CL j1 (R = N / CL of them; each row is its own 256-byte segment, so the stride is
//   free), transforms them locally and leaves W_N^(q k1) Y_q[k1] in its shared memory; after one cluster barrier
//   every CTA pulls the positions k1 in its 1/CL slice from all CL CTAs through distributed shared memory, does the
//   final radix-CL butterflies and stores X[k1 + R k2] straight to HBM. DSMEM moves (CL-1)/CL of a tile once
//   (B300_MICROARCH: ~20 B/cycle/SM, so it must stay below the HBM time of the tile). Shared memory is
//   [row][column] with the column fastest: every warp access is a run of whole 256-byte rows -- no padding, no
//   bank conflicts.
//
// Row passes (z axis, contiguous) -- fft_rows_r2c_kernel / fft_rows_c2r_kernel
//   r2c and c2r through a half-length complex transform with the split/merge step done in shared memory; one
//   buffer per row; the c2r results leave through a store functor so the particle emit can be fused into it.
#pragma once
#include <vector>
#include <cmath>
#include <type_traits>
#include <cooperative_groups.h>
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstdlib>
#include <cstring>
#include "gx_common.cuh"
#include "fft_core.cuh"

namespace gx {
namespace fft {

int ensure_twiddles();
template <typename T> const cplx<T> *twiddle_table(int n);
template <typename T> int fft3d(T *data, int n, int dir, cudaStream_t st);

// ---- element operators fused into the load or store side of a column pass -----------------------
// Called as op(value, i0, i1, i2) with the storage indices of the element in the [n0][n1][nc] array.
template <typename T> struct OpIdentity {
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int) const { return v; }
};
template <typename T> struct OpScale {
  T s;
  explicit OpScale(T s_) : s(s_) {}
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int) const { return mk<T>(v.x * s, v.y * s); }
};

// Store-side redirect (slab path, several GPUs): the inverse x-pass writes every element straight into the exchange buffer
// of the rank that owns its x plane -- peer-mapped memory, NVLink stores of whole TC-column row segments -- so the FFT
// transpose needs no send buffer, no all-to-all kernel and no second sweep: compute and collective are one kernel.
// Element (ix, ky_local, kz) of rank r lands in peer (ix / P) at [r][ix % P][ky_local][kz].
template <typename T> struct OpPushX {
  cplx<T> *peer[8];
  int pshift, pmask;       // P = n / world = 1 << pshift
  long self_off;           // r * P * P * nc: this rank's block in every peer's buffer
  long line_stride;        // P * nc: one x plane of a block
  static constexpr bool kPush = true;
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int) const { return v; }
  __device__ __forceinline__ cplx<T> *addr(int row, long col) const {
    return peer[row >> pshift] + (self_off + (long) (row & pmask) * line_stride + col);
  }
};
template <class S, class = void> struct is_push : std::false_type {};
template <class S> struct is_push<S, std::void_t<decltype(S::kPush)>> : std::true_type {};

// ---- which axis a column pass transforms ---------------------------------------------------------
// PassY: batch = i0, line index = i1, column = i2            (line stride nc,   batch stride n1*nc)
// PassX: one batch, line index = i0, column = i1*nc + i2     (line stride n1*nc)
struct PassY {};
struct PassX {};

// Where line position `row` of a batch lives: (row >> blk_shift) * blk_stride + (row & blk_mask) * line_stride.
// The plain layouts use one block (blk_shift = 30). The slab-decomposed y-pass reads (inverse) or writes (forward) the
// all-to-all exchange buffer directly -- [peer][x_local][ky_local][kz] -- so no pack/unpack sweep exists.
struct LineMap {
  long line_stride;    // elements between consecutive line positions inside a block
  long batch_stride;   // elements between batches
  long blk_stride;     // elements between blocks of line positions
  int blk_shift, blk_mask;
  __host__ __device__ __forceinline__ long off(int row) const {
    return (long) (row >> blk_shift) * blk_stride + (long) (row & blk_mask) * line_stride;
  }
};
inline LineMap plain_map(long line_stride, long batch_stride) {
  LineMap m; m.line_stride = line_stride; m.batch_stride = batch_stride; m.blk_stride = 0; m.blk_shift = 30; m.blk_mask = 0x3fffffff;
  return m;
}
inline LineMap blocked_map(long line_stride, long batch_stride, int rows_per_block, long blk_stride) {
  LineMap m; m.line_stride = line_stride; m.batch_stride = batch_stride; m.blk_stride = blk_stride;
  m.blk_shift = 0; while ((1 << m.blk_shift) < rows_per_block) ++m.blk_shift;
  m.blk_mask = rows_per_block - 1;
  return m;
}

struct ColGeo {
  int prefetch_tiles;  // > 0: every CTA also prefetches its rows of tile (tile + prefetch_tiles) into L2
  long ncols;          // columns per batch
  LineMap src, dst;
  int tiles_per_batch;
  int nc;              // fastest extent (n/2+1)
};

// default cluster size / tile width / residency target per (type, N); profiles/r1_fft_variants.txt keeps the sweep
template <typename T, int N> struct ColsPlan {
  // 128-byte row segments x 256 rows per CTA = 32 KB tiles, single buffer, one tile per cluster, five CTAs per SM
  // (register cap 102): on B200 the phase diversity of many small CTAs beat both the persistent double-buffered
  // form and wider tiles. 4096-point lines keep the portable cluster size 8 with 512 rows per CTA.
  static constexpr int TC = 128 / (int) sizeof(cplx<T>);
  static constexpr int Rmax = (N >= 4096) ? 512 : 256;
  static constexpr int CL = (N > Rmax) ? N / Rmax : 1;  // 512 -> 2, 1024 -> 4, 2048 -> 8, 4096 -> 8
  static constexpr int MINB = (N >= 4096) ? 1 : (sizeof(T) == 8 ? 5 : 4);
  static constexpr int SEL = 4;  // single buffer, non-persistent (see ColsCfg)
};

template <typename T, int N, int CL_, int TC_, int SEL_ = 0> struct ColsCfg {
  static constexpr int TC = TC_;
  static constexpr int CL = CL_;
  static constexpr int R = N / CL;
  // SEL_ bit 0: 8-points-per-thread plan (exists for 256-point local transforms); bit 2: single buffer, one tile per
  // cluster (non-persistent); bit 1: timing experiment only (wrong results): no DSMEM traffic
  // SEL_ bit 3: form the butterfly twiddles as powers of one table entry (as the row passes do)
  static constexpr int SEL = ((R == 256 && (SEL_ & 1)) ? 1 : 0) | ((SEL_ & 8) ? 2 : 0);
  static constexpr bool kExperimentLocalPull = (SEL_ & 2) != 0;
  static constexpr int NBUF = (SEL_ & 4) ? 1 : 2;
  using B = BlockFFT<T, R, -1, TC, SEL>;
  static constexpr int E = B::E, P = B::P;
  static constexpr int threads = P * TC;
  static constexpr int IPT = (CL > 1) ? E / CL : 1;  // cross-cluster butterflies per thread
  static_assert(CL == 1 || (E % CL == 0 && E >= CL), "first local radix must be a multiple of the cluster size");
  static_assert(threads <= 1024, "too many threads per CTA");
  static constexpr size_t tile_bytes = (size_t) R * TC * sizeof(cplx<T>);
  // two tile buffers + the local twiddles W_R^j and (clusters) this CTA's cross twiddles W_N^(q j)
  static constexpr size_t smem = NBUF * tile_bytes + (size_t) (CL > 1 ? 2 : 1) * R * sizeof(cplx<T>);
};

__device__ __forceinline__ unsigned smem_addr_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

// one element global -> shared without passing through registers (LDGSTS); !pred zero-fills
template <typename C> __device__ __forceinline__ void cp_async_elem(C *sdst, const C *gsrc, bool pred) {
  if constexpr (sizeof(C) == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr_u32(sdst)), "l"(gsrc), "r"(pred ? 16 : 0)
                 : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr_u32(sdst)), "l"(gsrc), "r"(pred ? 8 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NG> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NG) : "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// relaxed arrive that cannot be issued before `dep` has been computed (i.e. before the loads feeding it returned)
__device__ __forceinline__ void cluster_arrive_relaxed_after(double dep) {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::"d"(dep) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed_after(float dep) {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::"f"(dep) : "memory");
}

// Persistent: cluster `cid` of `nclusters` walks the tiles cid, cid + nclusters, ...; the loads of the next tile are in
// flight (cp.async into the other half of shared memory) while the current one is transformed.
// TMA = true: the CTA's tile -- R rows at stride CL lines x TC columns, 128-byte rows -- is fetched by ONE thread with a
// tiled tensor-map copy (cp.async.bulk.tensor, UTMALDG in SASS) completing on an mbarrier, instead of E cp.async per thread
// with their 64-bit address arithmetic. The tensor map views the matrix as [column][q = line mod CL][j = line / CL]([batch]), so
// the strided row set of CTA q is the dense box {TC, 1, R} at (column, q, 0). Columns past the end are zero-filled by the
// TMA unit. Used for plain (not exchange-layout) sources whose strides are multiples of 16 bytes.
template <typename T, int N, int CL, int TC, int MINB, int SEL, int DIR, class Pass, class LoadOp, class StoreOp, bool TMA = false>
__global__ void __launch_bounds__(ColsCfg<T, N, CL, TC, SEL>::threads, MINB)
fft_cols_kernel(const cplx<T> *src, cplx<T> *dst, ColGeo g, long ntiles, int nclusters, const cplx<T> *__restrict__ twR,
                const cplx<T> *__restrict__ twN, LoadOp lop, const __grid_constant__ StoreOp sop,
                const __grid_constant__ CUtensorMap tmap) {
  using C = cplx<T>;
  using Cfg = ColsCfg<T, N, CL, TC, SEL>;
  using B = BlockFFT<T, Cfg::R, DIR, TC, Cfg::SEL>;
  constexpr int R = Cfg::R, E = B::E, P = B::P;
  static_assert(!TMA || Cfg::NBUF == 1, "the TMA form is single-buffered");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  __shared__ __align__(8) unsigned long long tma_bar;
  unsigned tma_phase = 0;
  if constexpr (TMA) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(&tma_bar)), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }

  const int c = threadIdx.x % TC;
  const int t = threadIdx.x / TC;
  int q = 0;
  if constexpr (CL > 1) q = (int) cooperative_groups::this_cluster().block_rank();
  const long cid = blockIdx.x / CL;
  C *twl = smem + Cfg::NBUF * (R * TC);  // W_R^j
  C *twc = twl + R;              // W_N^(q j)
  // (Measured, profiles/r2s16_*: issuing the first tile's loads BEFORE this fill -- so that the L2 round trip of the twiddles
  // overlaps the tile's DRAM latency -- made the passes 4 % slower, 81 / 129 ms against 77 / 125: the fill is kept in front.)
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    twl[i] = twR[i];
    if constexpr (CL > 1) twc[i] = twN[q * i];
  }
  __syncthreads();

  auto issue = [&](long tile, C *buf) {
    const int batch = (int) (tile / g.tiles_per_batch);
    if constexpr (TMA) {
      if (threadIdx.x == 0) {
        constexpr int RB = R > 256 ? 256 : R;                       // a box dimension holds at most 256 elements
        const unsigned mb = smem_addr_u32(&tma_bar);
        const int c0 = (int) ((tile % g.tiles_per_batch) * TC) * 2;  // innermost dimension counts reals
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last touched through the generic proxy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned) (R * TC * sizeof(C))) : "memory");
#pragma unroll
        for (int j0 = 0; j0 < R; j0 += RB) {
          const unsigned dstp = smem_addr_u32(buf + (size_t) j0 * TC);
          if (std::is_same<Pass, PassY>::value)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dstp), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(c0), "r"(q), "r"(j0), "r"(batch), "r"(mb) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dstp), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(c0), "r"(q), "r"(j0), "r"(mb) : "memory");
        }
      }
      return;
    }
    const long col = (tile % g.tiles_per_batch) * TC + c;
    const bool active = col < g.ncols;
    const C *p = src + ((long) batch * g.src.batch_stride + (active ? col : 0));
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int row = (CL > 1) ? q + CL * (t + i * P) : t + i * P;
      cp_async_elem<C>(buf + (t + i * P) * TC + c, p + g.src.off(row), active);
    }
    cp_async_commit();
  };
  auto wait_tile = [&]() {
    if constexpr (TMA) {
      unsigned done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_addr_u32(&tma_bar)), "r"(tma_phase) : "memory");
      tma_phase ^= 1u;
    } else {
      cp_async_wait<0>();
    }
  };

  long tile = cid;
  if (Cfg::NBUF == 2 && tile < ntiles) issue(tile, smem);
  for (int it = 0; tile < ntiles; tile += nclusters, ++it) {
    C *buf = smem + (size_t) (Cfg::NBUF == 2 ? (it & 1) : 0) * (R * TC);
    if constexpr (Cfg::NBUF == 2) {
      const long next = tile + nclusters;
      if (next < ntiles) { issue(next, smem + (size_t) ((it + 1) & 1) * (R * TC)); cp_async_wait<1>(); }
      else cp_async_wait<0>();
    } else {
      issue(tile, buf);
      if (g.prefetch_tiles > 0) {
        // L2 prefetch of the rows this CTA position will load one wave of clusters later (GX_COLS_PREFETCH=<tiles>). Measured
        // at 2048^3 (profiles/r2s15_prefetch.txt): the x-pass goes from 76 to 92 ms and the y-pass from 124 to 128 at any
        // distance -- the hint competes with the demand loads instead of hiding their latency. Off by default.
        const long ptile = tile + g.prefetch_tiles;
        if (ptile < ntiles) {
          const int pbatch = (int) (ptile / g.tiles_per_batch);
          const long pcol = (ptile % g.tiles_per_batch) * TC;
          const C *pp = src + ((long) pbatch * g.src.batch_stride + pcol);
          for (int r = threadIdx.x; r < R; r += Cfg::threads) {
            const int row = (CL > 1) ? q + CL * r : r;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + g.src.off(row)));
          }
        }
      }
      wait_tile();
    }

    const int batch = (int) (tile / g.tiles_per_batch);
    const long col = (tile % g.tiles_per_batch) * TC + c;
    const bool active = col < g.ncols;
    int c1 = 0, c2;  // storage indices that do not depend on the line position
    if (std::is_same<Pass, PassY>::value) { c2 = (int) col; }
    else { c1 = (int) (col / g.nc); c2 = (int) (col - (long) c1 * g.nc); }
    const long base = (long) batch * g.dst.batch_stride + col;
    auto load_op = [&](C x, int row) -> C {
      return std::is_same<Pass, PassY>::value ? lop(x, batch, row, c2) : lop(x, row, c1, c2);
    };
    auto store = [&](int row, C val) {
      if (active) {
        C y = std::is_same<Pass, PassY>::value ? sop(val, batch, row, c2) : sop(val, row, c1, c2);
        if constexpr (is_push<StoreOp>::value) *sop.addr(row, base) = y;
        else dst[base + g.dst.off(row)] = y;
      }
    };

    // every thread consumes exactly the elements it copied itself, and stage 0 rewrites the same positions:
    // no barrier between the copy and the first butterflies
    C v[E];
    if constexpr (CL == 1) {
#pragma unroll
      for (int i = 0; i < E; ++i) v[i] = active ? load_op(buf[(t + i * P) * TC + c], t + i * P) : mk<T>(0, 0);
      B::run(v, t, buf + c, twl, [&](int k, C val) { store(k, val); });
      if constexpr (B::S > 1) __syncthreads();  // the buffer is the target of the copy issued two tiles later
    } else {
      // decimation in time across the cluster: CTA q owns the rows j = q + CL j1, transforms them locally and leaves
      // W_N^(q k1) Y_q[k1] in shared memory
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int row = q + CL * (t + i * P);
        v[i] = active ? load_op(buf[(t + i * P) * TC + c], row) : mk<T>(0, 0);
      }
      B::run_synced_emit(v, t, buf + c, twl, [&](int k1, C val) {
        buf[k1 * TC + c] = (q == 0) ? val : cmul(val, B::twiddle(twc, k1));
      });
      // release/acquire: the peers' DSMEM reads must see these shared-memory writes (a relaxed arrive after a CTA
      // barrier is NOT enough -- remote reads can overtake stores still queued in this SM; measured as sparse wrong
      // tiles). With a single buffer no global access is outstanding here, so the fence is cheap.
      cluster_arrive_release();
      cluster_wait();
      // final radix-CL stage: this CTA combines the positions k1 in [q R/CL, (q+1) R/CL) of all CL sub-transforms
      // (one pull through distributed shared memory) and stores X[k1 + R k2], k2 = 0..CL-1
      constexpr int IPT = Cfg::IPT;
      C y[IPT][CL];
      cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
#pragma unroll
      for (int m = 0; m < IPT; ++m) {
        const int k1 = q * (R / CL) + t + m * P;
#pragma unroll
        for (int q2 = 0; q2 < CL; ++q2)
          y[m][q2] = cluster.map_shared_rank(buf, Cfg::kExperimentLocalPull ? q : q2)[((k1 + (Cfg::kExperimentLocalPull ? q2 * 7 : 0)) % R) * TC + c];
      }
      T dep = (T) 0;
#pragma unroll
      for (int m = 0; m < IPT; ++m) {
        butterfly<CL, DIR, T>(y[m]);
        dep += y[m][0].x;  // output 0 is the sum of all CL pulled values
      }
      // Nobody may overwrite (next prefetch) or retire (exit) a buffer its peers are still pulling from: arrive only
      // once this thread's pulls have returned -- enforced by the data dependence, not by a memory barrier -- and
      // defer the wait past the stores.
      cluster_arrive_relaxed_after(dep);
#pragma unroll
      for (int m = 0; m < IPT; ++m) {
        const int k1 = q * (R / CL) + t + m * P;
#pragma unroll
        for (int k2 = 0; k2 < CL; ++k2) store(k1 + R * k2, y[m][k2]);
      }
      cluster_wait();
    }
  }
}

// Upper bound on the clusters a column pass launches (0: one per tile). A pass that is bound by NVLink rather than HBM (the
// slab x-pass that stores into its peers, OpPushX) needs only a few KB in flight per SM: launched as a small persistent grid on
// a side stream it leaves the SMs to the HBM-bound passes of the previous component running beside it.
inline int &cols_cluster_limit() { static int v = 0; return v; }

// ---- tensor maps for the TMA form of the column passes ---------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeFn) p;
    else
      cudaGetLastError();
  }
  return fn;
}

// [reals of a row of columns][q = line mod CL][j = line / CL]([batch]); false if the layout does not qualify
template <typename T, class Pass>
bool make_cols_tensor_map(CUtensorMap *tm, const cplx<T> *src, const ColGeo &g, int nbatch, int CL, int R, int TC) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc || g.src.blk_shift != 30 || ((size_t) src & 15)) return false;
  const size_t esz = sizeof(cplx<T>);
  const cuuint64_t line_b = (cuuint64_t) g.src.line_stride * esz, batch_b = (cuuint64_t) g.src.batch_stride * esz;
  if ((line_b & 15) || (std::is_same<Pass, PassY>::value && (batch_b & 15))) return false;
  const cuuint32_t rb = R > 256 ? 256 : R;
  cuuint64_t dims[4] = {(cuuint64_t) g.ncols * 2, (cuuint64_t) CL, (cuuint64_t) R, (cuuint64_t) nbatch};
  cuuint64_t strides[3] = {line_b, line_b * (cuuint64_t) CL, batch_b};
  cuuint32_t box[4] = {(cuuint32_t) TC * 2, 1, rb, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const int rank = std::is_same<Pass, PassY>::value ? 4 : 3;
  const CUresult rc = enc(tm, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t) rank,
                          (void *) src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

// Measured at 2048^3 f64 (profiles/r2s13_cols_tma.txt): x-pass 85.2 ms against 78.7 with cp.async, y-pass 133.9 against 126.7 --
// a box of 256 rows of 128 bytes is 256 short row fetches for the TMA unit, the same finding as round 1's per-row bulk copies.
// Correct (the GPU suite and the cuFFT cross-check pass with it), kept for wider tiles, off unless GX_COLS_TMA=1.
inline bool cols_tma_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("GX_COLS_TMA"); on = (e && e[0] == '1') ? 1 : 0; }
  return on != 0;
}

template <typename T, int N, int CL, int TC, int MINB, int SEL, int DIR, class Pass, class LoadOp, class StoreOp, bool TMA>
int launch_cols_variant(const cplx<T> *src, cplx<T> *dst, const ColGeo &g, long ntiles, LoadOp lop, StoreOp sop,
                        const CUtensorMap &tmap, cudaStream_t st) {
  using Cfg = ColsCfg<T, N, CL, TC, SEL>;
  auto kern = fft_cols_kernel<T, N, CL, TC, MINB, SEL, DIR, Pass, LoadOp, StoreOp, TMA>;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(Cfg::threads);
  cfg.dynamicSmemBytes = Cfg::smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // how many clusters can be resident at once (queried once per kernel and device)
  static int max_clusters = 0, max_clusters_dev = -1;
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  if (max_clusters_dev != dev) {
    if (Cfg::smem > 48 * 1024)
      GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
    // ask for the largest shared-memory carveout: residency is decided by how many double-buffered tiles fit
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (CL > 8) GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (CL > 1) {
      cfg.gridDim = dim3(CL * 1024);
      GX_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    } else {
      int per_sm = 0, sms = 0;
      GX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::threads, Cfg::smem));
      GX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      max_clusters = per_sm * sms;
    }
    if (max_clusters < 1) return fail("fft_cols: kernel cannot be resident (cluster occupancy query returned 0)");
    max_clusters_dev = dev;
    if (getenv("GX_DEBUG"))
      fprintf(stderr, "[gx] fft_cols<N=%d,CL=%d,TC=%d,%zu-byte elems,tma=%d>: %d resident clusters, %zu KB smem/CTA, %d threads\n", N,
              CL, TC, sizeof(cplx<T>), (int) TMA, max_clusters, Cfg::smem / 1024, Cfg::threads);
  }
  int nclusters = (int) ((ntiles < max_clusters || Cfg::NBUF == 1) ? ntiles : max_clusters);
  if (cols_cluster_limit() > 0 && nclusters > cols_cluster_limit()) nclusters = cols_cluster_limit();
  cfg.gridDim = dim3((unsigned) nclusters * CL);
  const char *tname = std::is_same<Pass, PassY>::value ? "fft_cols_y" : "fft_cols_x";
  GX_TIMED(tname, st);
  const cplx<T> *twR = twiddle_table<T>(Cfg::R), *twN = twiddle_table<T>(N);
  GX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, src, dst, g, ntiles, nclusters, twR, twN, lop, sop, tmap));
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

template <typename T, int N, int CL, int TC, int MINB, int SEL, int DIR, class Pass, class LoadOp, class StoreOp>
int launch_cols_cfg(const cplx<T> *src, cplx<T> *dst, ColGeo g, int nbatch, LoadOp lop, StoreOp sop, cudaStream_t st) {
  using Cfg = ColsCfg<T, N, CL, TC, SEL>;
  if constexpr (N == 4096 && CL == 8) {
    // 4096-point lines: the portable cluster size 8 means 512 rows per CTA -- 64 KB tiles, up to 1024 threads, one CTA per SM.
    // A (non-portable) cluster of 16 keeps the 2048^3 shape: 256 rows, 32 KB tiles, 4-5 CTAs per SM. GX_COLS4096_CL8=1: old form.
    static int cl16 = -1;
    if (cl16 < 0) { const char *e = getenv("GX_COLS4096_CL8"); cl16 = (e && e[0] == '1') ? 0 : 1; }
    if (cl16)
      return launch_cols_cfg<T, N, 16, TC, (sizeof(T) == 8 ? 5 : 4), SEL, DIR, Pass, LoadOp, StoreOp>(src, dst, g, nbatch, lop, sop, st);
  }
  g.tiles_per_batch = (int) ((g.ncols + TC - 1) / TC);
  const long ntiles = (long) g.tiles_per_batch * nbatch;
  if (ntiles <= 0) return 0;
  {
    static int dist = -2;  // GX_COLS_PREFETCH: prefetch distance in tiles (0: off)
    if (dist == -2) { const char *e = getenv("GX_COLS_PREFETCH"); dist = e ? atoi(e) : 0; }
    g.prefetch_tiles = (N >= 512 && g.src.blk_shift == 30) ? dist : 0;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if constexpr (N >= 512 && Cfg::NBUF == 1) {
    if (cols_tma_enabled() && make_cols_tensor_map<T, Pass>(&tmap, src, g, nbatch, CL, Cfg::R, TC))
      return launch_cols_variant<T, N, CL, TC, MINB, SEL, DIR, Pass, LoadOp, StoreOp, true>(src, dst, g, ntiles, lop, sop, tmap, st);
  }
  return launch_cols_variant<T, N, CL, TC, MINB, SEL, DIR, Pass, LoadOp, StoreOp, false>(src, dst, g, ntiles, lop, sop, tmap, st);
}

template <class Pass> inline ColGeo make_col_geo(int n0, int n1, int nc, int &nbatch) {
  ColGeo g;
  g.nc = nc;
  g.tiles_per_batch = 0;
  if (std::is_same<Pass, PassY>::value) { g.ncols = nc; g.src = plain_map(nc, (long) n1 * nc); nbatch = n0; }
  else { g.ncols = (long) n1 * nc; g.src = plain_map((long) n1 * nc, 0); nbatch = 1; }
  g.dst = g.src;
  return g;
}

// Column pass over an [n0][n1][nc] array: PassX transforms i0 (n0 == N), PassY transforms i1 (n1 == N).
template <typename T, int N, int DIR, class Pass, class LoadOp, class StoreOp>
int launch_cols(const cplx<T> *src, cplx<T> *dst, int n0, int n1, int nc, LoadOp lop, StoreOp sop, cudaStream_t st) {
  using PL = ColsPlan<T, N>;
  int nbatch;
  ColGeo g = make_col_geo<Pass>(n0, n1, nc, nbatch);
  return launch_cols_cfg<T, N, PL::CL, PL::TC, PL::MINB, PL::SEL, DIR, Pass, LoadOp, StoreOp>(src, dst, g, nbatch, lop, sop, st);
}

// ---- row passes (z axis, contiguous): r2c and c2r through a half-length complex FFT ---------------
template <typename T, int N> struct RowsCfg {
  static constexpr int M = N / 2;
  using B = BlockFFT<T, M, -1>;
  static constexpr int P = B::P;
  static constexpr int pitch = line_pitch<T>(M) + 1;  // >= M + 1 entries (the c2r input has M + 1)
  static constexpr size_t per_row = (size_t) pitch * sizeof(cplx<T>);
  // 128-thread CTAs (<= 37 KB of shared memory each) so that five are resident per SM in different phases
  static constexpr int byThreads = (128 / P) > 0 ? (128 / P) : 1;
  static constexpr int bySmem = (int) ((40 * 1024) / per_row) > 0 ? (int) ((40 * 1024) / per_row) : 1;
  static constexpr int LR = byThreads < bySmem ? byThreads : bySmem;
  static constexpr int threads = LR * P;
  static constexpr int MINB = (threads <= 128 && sizeof(T) == 8) ? 5 : 1;
  static constexpr size_t smem = (size_t) LR * per_row;
};

// Forward: each row of N reals (viewed as M = N/2 complex z[m] = x[2m] + i x[2m+1]) -> N/2+1 complex, in place.
//   X[k] = 1/2 (Z[k] + conj Z[M-k]) - i/2 W_N^k (Z[k] - conj Z[M-k]),  k = 0..M,  Z[M] == Z[0]
template <typename T, int N>
__global__ void __launch_bounds__(RowsCfg<T, N>::threads, RowsCfg<T, N>::MINB)
fft_rows_r2c_kernel(T *data, long nrows, const cplx<T> *__restrict__ twM, const cplx<T> *__restrict__ twN) {
  using C = cplx<T>;
  using Cfg = RowsCfg<T, N>;
  using B = BlockFFT<T, Cfg::M, -1>;
  constexpr int M = Cfg::M, E = B::E, P = B::P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  const int l = threadIdx.x / P, t = threadIdx.x % P;
  const long row = (long) blockIdx.x * Cfg::LR + l;
  const bool active = row < nrows;
  C *rowp = reinterpret_cast<C *>(data + row * (long) (N + 2));
  C *line = smem + (size_t) l * Cfg::pitch;
  C v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) v[i] = active ? rowp[t + i * P] : mk<T>(0, 0);
  // the spectrum Z lands in natural order in the same buffer the stages exchanged through
  B::run_synced_emit(v, t, line, twM, [&](int k, C val) { line[k] = val; });
  __syncthreads();
  if (!active) return;
  const T half = (T) 0.5;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int k = t + i * P;
    const C zk = line[k];
    const C zm = cconj(line[(M - k) & (M - 1)]);
    const C xe = mk<T>((zk.x + zm.x) * half, (zk.y + zm.y) * half);
    const C d = csub(zk, zm);
    const C xo = mk<T>(d.y * half, -d.x * half);  // -i/2 * d
    const C w = __ldg(twN + k);
    rowp[k] = cadd(xe, cmul(w, xo));
    if (k == 0) rowp[M] = mk<T>(zk.x - zk.y, (T) 0);  // X[M] = Re Z0 - Im Z0
  }
}

// c2r store functors: `begin(row)` is evaluated once per thread and row (everything that depends on the row only: base
// pointers, the x / y cell centre), `put(ctx, k, z)` once per complex result z = (x[2k], x[2k+1]), unscaled.
template <typename T> struct StoreRealScaled {
  T *data;
  T scale;
  int n;
  struct Ctx { cplx<T> *row; };
  __device__ __forceinline__ Ctx begin(long row) const {
    Ctx c; c.row = reinterpret_cast<cplx<T> *>(data + row * (long) (n + 2)); return c;
  }
  __device__ __forceinline__ void put(const Ctx &c, int k, cplx<T> z) const { c.row[k] = mk<T>(z.x * scale, z.y * scale); }
};

// Inverse: N/2+1 complex -> N reals (unnormalised c2r; the store functor scales).
//   Zt[k] = (X[k] + conj X[M-k]) + i W_N^{-k} (X[k] - conj X[M-k]),  z = IFFT_M(Zt),  x[2m] = Re z[m], x[2m+1] = Im z[m]
// The imaginary parts of X[0] and X[M] are ignored, as FFTW's c2r does.
// (Measured and dropped, profiles/r1s4_bench2048_table_{coalesce,nocoalesce}.json: routing the results through the line
// buffer so that a warp stores one contiguous run per instruction costs one more shared-memory round trip and barrier and
// was 14 % slower at 2048^3 -- the scattered 8..16-byte stores of the digit-reversed last stage are merged in L2 and are
// not what bounds this pass; the bytes in flight per SM are, hence the prefetching form below.)
template <typename T, int N, class Store>
__global__ void __launch_bounds__(RowsCfg<T, N>::threads, RowsCfg<T, N>::MINB)
fft_rows_c2r_kernel(const T *data, long row0, long nrows, const cplx<T> *__restrict__ twM,
                    const cplx<T> *__restrict__ twN, Store store) {
  using C = cplx<T>;
  using Cfg = RowsCfg<T, N>;
  using B = BlockFFT<T, Cfg::M, 1>;
  constexpr int M = Cfg::M, E = B::E, P = B::P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  const int l = threadIdx.x / P, t = threadIdx.x % P;
  const long row = row0 + (long) blockIdx.x * Cfg::LR + l;
  const bool active = row < row0 + nrows;
  const C *rowp = reinterpret_cast<const C *>(data + row * (long) (N + 2));
  C *line = smem + (size_t) l * Cfg::pitch;
#pragma unroll
  for (int i = 0; i < E; ++i) line[t + i * P] = active ? rowp[t + i * P] : mk<T>(0, 0);
  if (t == 0) line[M] = active ? rowp[M] : mk<T>(0, 0);
  __syncthreads();
  C v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int k = t + i * P;
    C xk = line[k];
    C xm = cconj(line[M - k]);
    if (k == 0) { xk.y = 0; xm.y = 0; }
    const C s = cadd(xk, xm);
    const C d = csub(xk, xm);
    const C w = cconj(__ldg(twN + k));
    const C wd = cmul(w, d);
    v[i] = mk<T>(s.x - wd.y, s.y + wd.x);  // s + i*wd
  }
  __syncthreads();  // everyone has consumed the inputs before the buffer is reused for the stage exchanges
  const auto ctx = store.begin(active ? row : row0);
  B::run(v, t, line, twM, [&](int k, C val) { if (active) store.put(ctx, k, val); });
}

// The same pass with its last stage decimation in time (fft_core.cuh DitRows): thread t emits k = t + (M/16) u, so the lanes
// of a warp store CONSECUTIVE results -- for the fused particle emit 256 contiguous bytes of pos and of vel per instruction
// instead of 32 scattered 8-byte stores.
template <typename T, int N> struct RowsDitCfg {
  static constexpr int M = N / 2;
  using D = DitRows<T, M, 1>;
  static constexpr int P = D::P;
  static constexpr int pitch = (M + 16 + 1 + 1) & ~1;  // raw row (M + 1 entries) and the 16 sub-transform rows (M + 16)
  static constexpr size_t per_row = (size_t) pitch * sizeof(cplx<T>);
  static constexpr int LR = (128 / P) > 0 ? (128 / P) : 1;
  static constexpr int threads = LR * P;
  static constexpr int MINB = (threads <= 128 && sizeof(T) == 8) ? 5 : (threads <= 128 ? 6 : 1);
  static constexpr size_t smem = (size_t) LR * per_row;
};

// TMA: a row of the padded field is (M + 1) complex = one contiguous, 16-byte aligned run in double precision (16400 bytes at
// 2048^3), so ONE thread stages the CTA's rows with cp.async.bulk (TMA bulk engine, UBLKCP in SASS) completing on an mbarrier,
// instead of 16 LDG + 16 STS per thread. (Single precision rows are 8-byte aligned only: they keep the register path.)
template <typename T, int N, class Store, bool TMA>
__global__ void __launch_bounds__(RowsDitCfg<T, N>::threads, RowsDitCfg<T, N>::MINB)
fft_rows_c2r_dit_kernel(const T *data, long row0, long nrows, const cplx<T> *__restrict__ twM,
                        const cplx<T> *__restrict__ twN, Store store, int prefetch_ctas) {
  using C = cplx<T>;
  using Cfg = RowsDitCfg<T, N>;
  using D = typename Cfg::D;
  constexpr int M = Cfg::M, P = Cfg::P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C *smem = reinterpret_cast<C *>(smem_raw);
  const int l = threadIdx.x / P, t = threadIdx.x % P;
  const long first = row0 + (long) blockIdx.x * Cfg::LR;
  const long row = first + l;
  const bool active = row < row0 + nrows;
  const C *rowp = reinterpret_cast<const C *>(data + row * (long) (N + 2));
  C *line = smem + (size_t) l * Cfg::pitch;
  if constexpr (TMA) {
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned mb = smem_addr_u32(&mbar);
    constexpr unsigned row_bytes = (unsigned) ((M + 1) * sizeof(C));
    static_assert(row_bytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (threadIdx.x == 0) {  // initialise and issue at once: the other threads only need the barrier when they wait on it
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const long left = row0 + nrows - first;
      const int nact = (int) (left < (long) Cfg::LR ? left : (long) Cfg::LR);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(row_bytes * (unsigned) nact) : "memory");
      for (int r = 0; r < nact; ++r)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr_u32(smem + (size_t) r * Cfg::pitch)),
                     "l"(data + (first + r) * (long) (N + 2)), "r"(row_bytes), "r"(mb)
                     : "memory");
      if (prefetch_ctas > 0) {  // the rows of the CTA that will run about one wave later, into L2
        const long pf = first + (long) prefetch_ctas * Cfg::LR;
        for (int r = 0; r < Cfg::LR; ++r)
          if (pf + r < row0 + nrows)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(data + (pf + r) * (long) (N + 2)), "r"(row_bytes) : "memory");
      }
    }
    if (!active) {
#pragma unroll
      for (int i = 0; i < 16; ++i) line[t + i * P] = mk<T>(0, 0);
      if (t == 0) line[M] = mk<T>(0, 0);
    }
    __syncthreads();  // the initialised barrier is visible to every thread
    unsigned done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mb), "r"(0) : "memory");
    __syncthreads();
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) line[t + i * P] = active ? rowp[t + i * P] : mk<T>(0, 0);
    if (t == 0) line[M] = active ? rowp[M] : mk<T>(0, 0);
    __syncthreads();
  }
  C v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int k = D::input_index(t, e);
    C xk = line[k];
    C xm = cconj(line[M - k]);
    if (k == 0) { xk.y = 0; xm.y = 0; }
    const C s = cadd(xk, xm);
    const C d = csub(xk, xm);
    const C w = cconj(__ldg(twN + k));
    const C wd = cmul(w, d);
    v[e] = mk<T>(s.x - wd.y, s.y + wd.x);  // s + i*wd
  }
  __syncthreads();  // everyone has consumed the inputs before the buffer is reused for the stage exchanges
  D::stage_a(v, t, line, twM);
  __syncthreads();
  if constexpr (D::RB > 1) {
    D::stage_b(t, line);
    __syncthreads();
  }
  const auto ctx = store.begin(active ? row : row0);
  D::stage_final(t, line, twM, [&](int k, C val) { if (active) store.put(ctx, k, val); });
}

// (Also measured and dropped, profiles/r1s5_bench2048_rows_{single,pipelined}.json: a persistent form of this kernel with
// the next group of rows prefetched by cp.async into a second shared-memory buffer -- 3 CTAs of 128 threads per SM instead
// of 5 -- took 203 ms per step at 2048^3 against 174 ms: the pass is bound by how many warps an SM has to cover the
// shared-memory exchange and FP64 latency of the radix stages, not by loads in flight.)

template <typename T, int N> int launch_rows_r2c(T *data, long nrows, cudaStream_t st) {
  using Cfg = RowsCfg<T, N>;
  auto kern = fft_rows_r2c_kernel<T, N>;
  if (Cfg::smem > 48 * 1024)
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
  const long grid = (nrows + Cfg::LR - 1) / Cfg::LR;
  if (grid <= 0) return 0;
  GX_TIMED("fft_rows_r2c", st);
  kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, nrows, twiddle_table<T>(N / 2), twiddle_table<T>(N));
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

// rows [row0, row0 + nrows) of `data` (pitch N+2 reals) -> store functor
inline bool rows_dit_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("GX_ROWS_DIT"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

template <typename T, int N, class Store>
int launch_rows_c2r(const T *data, long row0, long nrows, Store store, cudaStream_t st) {
  if constexpr (N >= 512 && N <= 4096) {
    if (rows_dit_enabled()) {
      using Cfg = RowsDitCfg<T, N>;
      static int tma = -1;  // GX_ROWS_TMA=0: stage the rows through registers (A/B switch)
      if (tma < 0) { const char *e = getenv("GX_ROWS_TMA"); tma = (e && e[0] == '0') ? 0 : 1; }
      const bool use_tma = sizeof(T) == 8 && tma && ((size_t) data % 16 == 0);
      static int pf = -2;   // GX_ROWS_PREFETCH: L2 prefetch distance in CTAs (0: off)
      if (pf == -2) { const char *e = getenv("GX_ROWS_PREFETCH"); pf = e ? atoi(e) : 0; }
      const long grid = (nrows + Cfg::LR - 1) / Cfg::LR;
      if (grid <= 0) return 0;
      GX_TIMED("fft_rows_c2r", st);
      if constexpr (sizeof(T) == 8) {
        if (use_tma) {
          auto kern = fft_rows_c2r_dit_kernel<T, N, Store, true>;
          if (Cfg::smem > 48 * 1024)
            GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
          kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, row0, nrows, twiddle_table<T>(N / 2), twiddle_table<T>(N), store, pf);
          count_launch();
          GX_LAUNCH_CHECK();
          return 0;
        }
      }
      auto kern = fft_rows_c2r_dit_kernel<T, N, Store, false>;
      if (Cfg::smem > 48 * 1024)
        GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
      kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, row0, nrows, twiddle_table<T>(N / 2), twiddle_table<T>(N), store, 0);
      count_launch();
      GX_LAUNCH_CHECK();
      return 0;
    }
  }
  using Cfg = RowsCfg<T, N>;
  auto kern = fft_rows_c2r_kernel<T, N, Store>;
  if (Cfg::smem > 48 * 1024)
    GX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::smem));
  const long grid = (nrows + Cfg::LR - 1) / Cfg::LR;
  if (grid <= 0) return 0;
  GX_TIMED("fft_rows_c2r", st);
  kern<<<(unsigned) grid, Cfg::threads, Cfg::smem, st>>>(data, row0, nrows, twiddle_table<T>(N / 2),
                                                        twiddle_table<T>(N), store);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

}  // namespace fft
}  // namespace gx

#define GX_N_SWITCH(n, CALL)                                                                                     \
  switch (n) {                                                                                                   \
    case 4: { constexpr int N = 4; return CALL; }                                                                \
    case 8: { constexpr int N = 8; return CALL; }                                                                \
    case 16: { constexpr int N = 16; return CALL; }                                                              \
    case 32: { constexpr int N = 32; return CALL; }                                                              \
    case 64: { constexpr int N = 64; return CALL; }                                                              \
    case 128: { constexpr int N = 128; return CALL; }                                                            \
    case 256: { constexpr int N = 256; return CALL; }                                                            \
    case 512: { constexpr int N = 512; return CALL; }                                                            \
    case 1024: { constexpr int N = 1024; return CALL; }                                                          \
    case 2048: { constexpr int N = 2048; return CALL; }                                                          \
    case 4096: { constexpr int N = 4096; return CALL; }                                                          \
    default: return fail("grid size must be a power of two in [4,4096]");                                        \
  }


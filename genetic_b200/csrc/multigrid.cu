// Moving real-space fields between grids of different resolution (zoom levels), windows and flagged-cell masks.
//
// Reference (paths relative to genetIC/src/):
//   Field::addFieldFromDifferentGrid             simulation/field/field.hpp:804-812  (through evaluator.hpp:31-57)
//     coarse -> fine : SuperSampleEvaluator[SectionEvaluator] -> Field::evaluateInterpolated :483-508,
//                      makeTricubicInterpolator :598-617, LocalUnitTricubicApproximation tools/numerics/tricubic.hpp:54-77
//     fine -> coarse : SubSampleEvaluator evaluator.hpp:168-205, SubSampleGrid::forEachSubcell virtualgrid.hpp:570-600
//   Field::deInterpolate                          field.hpp:420-480 in the loop of
//     MultiLevelGridBase::generateMultilevelCovectorFromHiresCovector multilevelgrid.hpp:283-297
//   Field::setZeroInsideWindow / setZeroOutsideWindow  field.hpp:774-795, Window<T>::contains window.hpp
//   FilteredVarianceModification::windowOperator / varianceOperator quadraticmodification.hpp:172-221
//
// All three transfers are separable: out[A][B][C] (+)= scale * sum_ijk wx[A][i] wy[B][j] wz[C][k] v[ix[A][i]][iy[B][j]][iz[C][k]].
// The tricubic polynomial of the reference (64 coefficients from a 4^3 neighbourhood, tricubic.hpp:80-430) is the tensor
// product of the per-axis Catmull-Rom weights listed at tricubic.hpp:450-549, so interpolation gathers with those
// weights, "conjugate deinterpolation" gathers with the transposed per-axis tables, and sub-cell averaging gathers with
// unit weights and divides by the number of sub-cells the fine grid holds. The per-axis tables (a few KB) are built on
// the host with the reference's own coordinate arithmetic (fmod wrap, floor(x/dx - 1/2)) in the field's CoordinateType;
// the kernel is one CTA per destination row: the 4x4 (or (4r)^2, r^2) source rows are combined into one row in shared
// memory with coalesced loads, then every thread finishes one cell along z.
#include "../../include/genetic_b200.h"
#include "gx_common.cuh"
#include "spectral.cuh"
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace gx {

template <typename T> struct AxisTab {
  const int *cnt;            // [n_dst]
  const int *idx;            // [n_dst][K] source indices
  const T *w;                // [n_dst][K]
  const unsigned char *flag; // [n_dst] bit 0 / bit 1: containment tests
  int K;
};

// mode 0: weighted sum, cell contained when bit 0 is set on all three axes
// mode 1: mean over the gathered cells, contained when bit 0 is set on all axes or bit 1 is set on all axes
// One CTA makes `group` consecutive destination lines of one x-plane: their source lines (cx * cy each, 4 x 4 for the tricubic
// transfer) overlap, and the CTA's later lines find them in L1 instead of L2.
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
axis_gather_add_kernel(T *__restrict__ dst, int n_dst, const T *__restrict__ src, int n_src, AxisTab<T> tx, AxisTab<T> ty,
                       AxisTab<T> tz, T scale, int group) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *row = reinterpret_cast<T *>(smem_raw);
  const int A = blockIdx.y;
  const int cx = tx.cnt[A];
  if (cx == 0) return;
  const long ps = n_src + 2, pd = n_dst + 2;
  const int *ix = tx.idx + (long) A * tx.K;
  const T *wx = tx.w + (long) A * tx.K;
  const int B1 = min(n_dst, (int) (blockIdx.x + 1) * group);
  for (int B = blockIdx.x * group; B < B1; ++B) {  // every condition below is uniform over the CTA
    const int cy = ty.cnt[B];
    if (cy == 0) continue;
    const unsigned fxy = tx.flag[A] & ty.flag[B];
    if (fxy == 0) continue;
    const int *iy = ty.idx + (long) B * ty.K;
    const T *wy = ty.w + (long) B * ty.K;
    for (int z = threadIdx.x; z < n_src; z += blockDim.x) {
      T acc = 0;
      for (int i = 0; i < cx; ++i) {
        const T *plane = src + (long) ix[i] * n_src * ps + z;
        T a = 0;
        for (int j = 0; j < cy; ++j) a += wy[j] * plane[(long) iy[j] * ps];
        acc += wx[i] * a;
      }
      row[z] = acc;
    }
    __syncthreads();
    T *out = dst + ((long) A * n_dst + B) * pd;
    for (int C = threadIdx.x; C < n_dst; C += blockDim.x) {
      const int cz = tz.cnt[C];
      if (cz == 0) continue;
      const unsigned f = fxy & tz.flag[C];
      if (MODE == 0 ? !(f & 1u) : f == 0) continue;
      const int *iz = tz.idx + (long) C * tz.K;
      const T *wz = tz.w + (long) C * tz.K;
      T s = 0;
      for (int k = 0; k < cz; ++k) s += wz[k] * row[iz[k]];
      if (MODE == 1) s /= (T) (cx * cy * cz);
      out[C] += scale * s;
    }
    __syncthreads();
  }
}

// The same gather with at most K source cells per axis (the tricubic transfer has 4, a factor-2 sub-cell mean 2): the
// per-axis tables of the CTA's x index, of the current destination line and of the thread's z cells sit in registers, so the
// inner loops are the K x K data loads and FMAs only -- the generic kernel above re-reads index and weight for every term
// (48 load instructions per output for K = 4; the load/store unit, not HBM, was its bound: 0.42 ms per 256^3 transfer
// against 0.05 ms of traffic).
template <typename T, int MODE, int K>
__global__ void __launch_bounds__(256)
axis_gather_add_small_kernel(T *__restrict__ dst, int n_dst, const T *__restrict__ src, int n_src, AxisTab<T> tx, AxisTab<T> ty,
                             AxisTab<T> tz, T scale, int group) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *row = reinterpret_cast<T *>(smem_raw);
  const int A = blockIdx.y;
  const int cx = tx.cnt[A];
  if (cx == 0) return;
  const long ps = n_src + 2, pd = n_dst + 2;
  long xoff[K];
  T wx[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const bool on = i < cx;
    xoff[i] = on ? (long) tx.idx[(long) A * tx.K + i] * n_src * ps : 0;
    wx[i] = on ? tx.w[(long) A * tx.K + i] : (T) 0;
  }
  const unsigned fx = tx.flag[A];
  // z tables of the (up to two) destination cells this thread finishes: n_dst <= 2 * blockDim.x here
  int zc[2], zi[2][K];
  T zw[2][K];
  unsigned zf[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int C = threadIdx.x + u * blockDim.x;
    zc[u] = C < n_dst ? tz.cnt[C] : 0;
    zf[u] = C < n_dst ? tz.flag[C] : 0u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool on = k < zc[u];
      zi[u][k] = on ? tz.idx[(long) C * tz.K + k] : 0;
      zw[u][k] = on ? tz.w[(long) C * tz.K + k] : (T) 0;
    }
  }
  const int B1 = min(n_dst, (int) (blockIdx.x + 1) * group);
  for (int B = blockIdx.x * group; B < B1; ++B) {  // every condition on A, B is uniform over the CTA
    const int cy = ty.cnt[B];
    const unsigned fxy = fx & ty.flag[B];
    if (cy == 0 || fxy == 0) continue;
    long yoff[K];
    T wy[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const bool on = j < cy;
      yoff[j] = on ? (long) ty.idx[(long) B * ty.K + j] * ps : 0;
      wy[j] = on ? ty.w[(long) B * ty.K + j] : (T) 0;
    }
    T *out = dst + ((long) A * n_dst + B) * pd;
    T prev[2];  // the read of the read-modify-write goes out with the source loads
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int C = threadIdx.x + u * blockDim.x;
      prev[u] = (C < n_dst && zc[u] != 0) ? out[C] : (T) 0;
    }
    for (int z = threadIdx.x; z < n_src; z += blockDim.x) {
      T acc = 0;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (i < cx) {
          const T *plane = src + xoff[i] + z;
          T a = 0;
#pragma unroll
          for (int j = 0; j < K; ++j)
            if (j < cy) a += wy[j] * plane[yoff[j]];
          acc += wx[i] * a;
        }
      }
      row[z] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int C = threadIdx.x + u * blockDim.x;
      if (C >= n_dst || zc[u] == 0) continue;
      const unsigned f = fxy & zf[u];
      if (MODE == 0 ? !(f & 1u) : f == 0) continue;
      T s = 0;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k < zc[u]) s += zw[u][k] * row[zi[u][k]];
      if (MODE == 1) s /= (T) (cx * cy * zc[u]);
      out[C] = prev[u] + scale * s;
    }
    __syncthreads();
  }
}

// ---- host-side per-axis tables ---------------------------------------------------------------------------------
template <typename T> struct HostAxis {
  int n = 0, K = 0;
  std::vector<int> cnt, idx;
  std::vector<T> w;
  std::vector<unsigned char> flag;
  void init(int n_, int K_) { n = n_; K = K_; cnt.assign(n, 0); idx.assign((size_t) n * K, 0); w.assign((size_t) n * K, (T) 0); flag.assign(n, 0); }
};

template <typename T> static inline T wrap_point(T x, T L) {  // Grid::wrapIndividualCoordinate, grid.hpp:527-531
  x = std::fmod(x, L);
  if (x < 0) x += L;
  return x;
}
template <typename T> static inline T centroid(const gx_grid_desc &g, int axis, int c) {  // grid.hpp:662-669
  return ((T) c * (T) g.cell_size + (T) g.offset_lower[axis]) + (T) g.cell_size / 2;
}
static inline int sim_equivalent(const gx_grid_desc &g) { return (int) std::lround(g.domain_size / g.cell_size); }  // grid.hpp:111

template <typename T> static inline void catmull_rom(T t, T w[4]) {  // tricubic.hpp:450-549, per axis
  w[0] = -(t - 1) * (t - 1) * t / 2;
  w[1] = (2 - 5 * t * t + 3 * t * t * t) / 2;
  w[2] = -t * (-1 - 4 * t + 3 * t * t) / 2;
  w[3] = (t - 1) * t * t / 2;
}

// the four source cells and weights Field::evaluateInterpolated uses for a point at the centroid of dst cell c
template <typename T>
static void stencil(const gx_grid_desc &dst, const gx_grid_desc &src, int axis, int c, bool wrap_with_grid, int idx[4], T w[4]) {
  T loc = centroid<T>(dst, axis, c) - (T) src.offset_lower[axis];
  loc = wrap_point<T>(loc, (T) src.domain_size);
  const T u = loc / (T) src.cell_size - (T) 0.5;
  const int p0 = (int) std::floor(u);
  catmull_rom<T>(u - (T) p0, w);
  const int S = sim_equivalent(src);
  for (int i = 0; i < 4; ++i) {
    int q = p0 - 1 + i;
    if (wrap_with_grid) {            // deInterpolate's own wrap (field.hpp:452-462), with >= where the reference has >
      if (q < 0) q += src.n;
      if (q >= src.n) q -= src.n;
    } else if (src.n == S) {         // wrapCoordinate, grid.hpp:566-574
      if (q > S - 1) q -= S;
      if (q < 0) q += S;
    } else {                         // clampCoordinate, grid.hpp:592-600
      if (q > src.n - 1) q = src.n - 1;
      if (q < 0) q = 0;
    }
    idx[i] = q;
  }
}

template <typename T> static int build_interp(const gx_grid_desc &dst, const gx_grid_desc &src, HostAxis<T> t[3]) {
  const int ratio = (int) std::lround(src.cell_size / dst.cell_size);
  if (ratio < 1 || std::fabs(src.cell_size / dst.cell_size - ratio) > 1e-6 * ratio)
    return fail("gx_interp_add: cell sizes are not in an integer ratio");
  const long S = (long) sim_equivalent(src) * ratio;
  for (int a = 0; a < 3; ++a) {
    t[a].init(dst.n, 4);
    const long rel = std::lround((dst.offset_lower[a] - src.offset_lower[a]) / dst.cell_size);
    for (int c = 0; c < dst.n; ++c) {
      int idx[4]; T w[4];
      stencil<T>(dst, src, a, c, false, idx, w);
      t[a].cnt[c] = 4;
      for (int i = 0; i < 4; ++i) { t[a].idx[(size_t) c * 4 + i] = idx[i]; t[a].w[(size_t) c * 4 + i] = w[i]; }
      long q = c + rel;  // SectionOfGrid / SuperSampleGrid containment, virtualgrid.hpp:199-206,472-480
      if (q > S - 1) q -= S;
      if (q < 0) q += S;
      t[a].flag[c] = (q >= 0 && q / ratio < src.n) ? 1 : 0;
    }
  }
  return 0;
}

template <typename T> static int build_deinterp(const gx_grid_desc &lo, const gx_grid_desc &hi, HostAxis<T> t[3]) {
  const int ratio = (int) std::lround(lo.cell_size / hi.cell_size);
  if (ratio < 1) return fail("gx_deinterp_add: the source must be the finer grid");
  for (int a = 0; a < 3; ++a) {
    std::vector<std::vector<std::pair<int, T>>> lists(lo.n);
    for (int f = 0; f < hi.n; ++f) {
      int idx[4]; T w[4];
      stencil<T>(hi, lo, a, f, true, idx, w);
      for (int i = 0; i < 4; ++i) lists[idx[i]].push_back({f, w[i]});
    }
    size_t K = 1;
    for (auto &l : lists) K = std::max(K, l.size());
    t[a].init(lo.n, (int) K);
    for (int p = 0; p < lo.n; ++p) {
      t[a].cnt[p] = (int) lists[p].size();
      t[a].flag[p] = 1;
      for (size_t i = 0; i < lists[p].size(); ++i) {
        t[a].idx[(size_t) p * K + i] = lists[p][i].first;
        t[a].w[(size_t) p * K + i] = lists[p][i].second;
      }
    }
  }
  return 0;
}

template <typename T> static int build_subsample(const gx_grid_desc &dst, const gx_grid_desc &src, HostAxis<T> t[3]) {
  const int r = (int) std::lround(dst.cell_size / src.cell_size);
  if (r < 2) return fail("gx_interp_add: sub-cell averaging needs a finer source");
  const long S = (long) sim_equivalent(dst) * r;
  for (int a = 0; a < 3; ++a) {
    t[a].init(dst.n, r);
    const long rel = std::lround((dst.offset_lower[a] - src.offset_lower[a]) / src.cell_size);
    for (int c = 0; c < dst.n; ++c) {
      int cnt = 0;
      unsigned char flag = 0;
      for (int j = 0; j < r; ++j) {
        long f = (long) c * r + j + rel;
        if (f > S - 1) f -= S;
        if (f < 0) f += S;
        if (f < src.n) {
          t[a].idx[(size_t) c * r + cnt] = (int) f;
          t[a].w[(size_t) c * r + cnt] = (T) 1;
          ++cnt;
          if (j == 0) flag |= 1;      // first sub-cell held  (virtualgrid.hpp:628-633)
          if (j == r - 1) flag |= 2;  // last sub-cell held
        }
      }
      t[a].cnt[c] = cnt;
      t[a].flag[c] = flag;
    }
  }
  return 0;
}

// device copies of the tables of one call; freed after the kernel has been enqueued (stream-ordered)
template <typename T> struct DevTabs {
  void *blob = nullptr;
  AxisTab<T> tab[3];
  int upload(const HostAxis<T> h[3], cudaStream_t st) {
    size_t off[3][4], total = 0;
    auto al = [](size_t x) { return (x + 15) & ~(size_t) 15; };
    for (int a = 0; a < 3; ++a) {
      off[a][0] = total; total += al(h[a].cnt.size() * sizeof(int));
      off[a][1] = total; total += al(h[a].idx.size() * sizeof(int));
      off[a][2] = total; total += al(h[a].w.size() * sizeof(T));
      off[a][3] = total; total += al(h[a].flag.size());
    }
    std::vector<unsigned char> stage(total);
    for (int a = 0; a < 3; ++a) {
      memcpy(&stage[off[a][0]], h[a].cnt.data(), h[a].cnt.size() * sizeof(int));
      memcpy(&stage[off[a][1]], h[a].idx.data(), h[a].idx.size() * sizeof(int));
      memcpy(&stage[off[a][2]], h[a].w.data(), h[a].w.size() * sizeof(T));
      memcpy(&stage[off[a][3]], h[a].flag.data(), h[a].flag.size());
    }
    GX_CUDA_CHECK(cudaMalloc(&blob, total));
    GX_CUDA_CHECK(cudaMemcpy(blob, stage.data(), total, cudaMemcpyHostToDevice));  // synchronous: `stage` is a local
    (void) st;
    unsigned char *b = (unsigned char *) blob;
    for (int a = 0; a < 3; ++a) {
      tab[a].cnt = (const int *) (b + off[a][0]);
      tab[a].idx = (const int *) (b + off[a][1]);
      tab[a].w = (const T *) (b + off[a][2]);
      tab[a].flag = b + off[a][3];
      tab[a].K = h[a].K;
    }
    return 0;
  }
};

// The tables depend only on (kind, dtype, the two grid descriptors): the metric, the covector construction and the variance
// iterations repeat the same few transfers hundreds of times, so the device copies are kept (a few KB each, at most 64
// entries, dropped oldest-first) instead of being rebuilt, allocated, copied and freed -- with a stream sync -- per call.
struct TabKey {
  int kind, dtype, device;
  gx_grid_desc a, b;
  bool operator==(const TabKey &o) const {
    return kind == o.kind && dtype == o.dtype && device == o.device && memcmp(&a, &o.a, sizeof(a)) == 0 && memcmp(&b, &o.b, sizeof(b)) == 0;
  }
};
struct TabEntry { TabKey key; void *tabs; };  // tabs: heap DevTabs<T> of the entry's dtype
static std::vector<TabEntry> g_tab_cache;

static TabKey make_key(int kind, int dtype, const gx_grid_desc &a, const gx_grid_desc &b) {
  TabKey k;
  memset(&k, 0, sizeof(k));  // descriptors are compared bytewise: clear the padding
  k.kind = kind; k.dtype = dtype;
  cudaGetDevice(&k.device);
  k.a.n = a.n; k.a.cell_size = a.cell_size; k.a.domain_size = a.domain_size;
  k.b.n = b.n; k.b.cell_size = b.cell_size; k.b.domain_size = b.domain_size;
  for (int i = 0; i < 3; ++i) { k.a.offset_lower[i] = a.offset_lower[i]; k.b.offset_lower[i] = b.offset_lower[i]; }
  return k;
}

void release_multigrid_workspace() {
  if (g_tab_cache.empty()) return;
  cudaDeviceSynchronize();
  for (auto &e : g_tab_cache) {
    cudaFree(((DevTabs<double> *) e.tabs)->blob);  // blob is the first member for either element type
    if (e.key.dtype == GX_F64) delete (DevTabs<double> *) e.tabs; else delete (DevTabs<float> *) e.tabs;
  }
  g_tab_cache.clear();
}

// kind 0: interpolate (coarse -> fine), 1: sub-cell mean (fine -> coarse), 2: deinterpolate. Builds on a miss.
template <typename T, class Build>
static int cached_tabs(int kind, const gx_grid_desc &a, const gx_grid_desc &b, Build build, cudaStream_t st, DevTabs<T> **out) {
  const TabKey key = make_key(kind, sizeof(T) == 8 ? GX_F64 : GX_F32, a, b);
  for (auto &e : g_tab_cache)
    if (e.key == key) { *out = (DevTabs<T> *) e.tabs; return 0; }
  HostAxis<T> h[3];
  if (build(h)) return 1;
  DevTabs<T> *d = new DevTabs<T>();
  if (d->upload(h, st)) { delete d; return 1; }
  if (g_tab_cache.size() >= 64) {
    GX_CUDA_CHECK(cudaDeviceSynchronize());  // a kernel may still be reading the oldest entry
    TabEntry old = g_tab_cache.front();
    g_tab_cache.erase(g_tab_cache.begin());
    cudaFree(((DevTabs<double> *) old.tabs)->blob);  // blob is the first member for either element type
    if (old.key.dtype == GX_F64) delete (DevTabs<double> *) old.tabs; else delete (DevTabs<float> *) old.tabs;
  }
  g_tab_cache.push_back(TabEntry{key, d});
  *out = d;
  return 0;
}

template <typename T, int MODE>
static int run_gather(void *dst, int n_dst, const void *src, int n_src, const DevTabs<T> &d, T scale, const char *name,
                      cudaStream_t st) {
  GX_TIMED(name, st);
  static const int group = [] { const char *e = getenv("GX_GATHER_GROUP"); const int g = e ? atoi(e) : 8; return g < 1 ? 1 : g; }();
  static const bool small = [] { const char *e = getenv("GX_GATHER_SMALL"); return !(e && e[0] == '0'); }();
  const int kmax = d.tab[0].K > d.tab[1].K ? (d.tab[0].K > d.tab[2].K ? d.tab[0].K : d.tab[2].K) : (d.tab[1].K > d.tab[2].K ? d.tab[1].K : d.tab[2].K);
  if (small && kmax <= 4 && n_dst <= 512) {  // register-resident tables (tricubic transfer, factor-2..4 sub-cell means)
    const dim3 grid((n_dst + group - 1) / group, n_dst);
    if (kmax <= 2)
      axis_gather_add_small_kernel<T, MODE, 2><<<grid, 256, (size_t) n_src * sizeof(T), st>>>(
          (T *) dst, n_dst, (const T *) src, n_src, d.tab[0], d.tab[1], d.tab[2], scale, group);
    else
      axis_gather_add_small_kernel<T, MODE, 4><<<grid, 256, (size_t) n_src * sizeof(T), st>>>(
          (T *) dst, n_dst, (const T *) src, n_src, d.tab[0], d.tab[1], d.tab[2], scale, group);
    count_launch();
    GX_LAUNCH_CHECK();
    return 0;
  }
  axis_gather_add_kernel<T, MODE><<<dim3((n_dst + group - 1) / group, n_dst), 256, (size_t) n_src * sizeof(T), st>>>(
      (T *) dst, n_dst, (const T *) src, n_src, d.tab[0], d.tab[1], d.tab[2], scale, group);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

static int check_desc(const gx_grid_desc *g, const char *who) {
  if (!g) return fail(std::string(who) + ": null grid descriptor");
  // real-space gathers with per-axis tables: any even size (virtual super/subsampled grids of the particle mappers are
  // e.g. 3 x 32 = 96 cells); only the transforms need powers of two
  if (g->n < 4 || g->n > 4096 || (g->n & 1)) return fail(std::string(who) + ": grid size must be even and in [4,4096]");
  if (!(g->cell_size > 0) || !(g->domain_size > 0)) return fail(std::string(who) + ": cell and domain size must be positive");
  return 0;
}

template <typename T>
static int interp_add_t(void *dst, const gx_grid_desc *gd, const void *src, const gx_grid_desc *gs, cudaStream_t st) {
  DevTabs<T> *d = nullptr;
  if (gs->cell_size > gd->cell_size) {
    if (cached_tabs<T>(0, *gd, *gs, [&](HostAxis<T> *h) { return build_interp<T>(*gd, *gs, h); }, st, &d)) return 1;
    return run_gather<T, 0>(dst, gd->n, src, gs->n, *d, (T) 1, "interp_add", st);
  }
  if (gs->cell_size < gd->cell_size) {
    if (cached_tabs<T>(1, *gd, *gs, [&](HostAxis<T> *h) { return build_subsample<T>(*gd, *gs, h); }, st, &d)) return 1;
    return run_gather<T, 1>(dst, gd->n, src, gs->n, *d, (T) 1, "subsample_add", st);
  }
  return fail("gx_interp_add: source and destination have the same resolution");
}

template <typename T>
static int deinterp_add_t(void *dst, const gx_grid_desc *gc, const void *src, const gx_grid_desc *gf, double scale,
                          cudaStream_t st) {
  DevTabs<T> *d = nullptr;
  if (cached_tabs<T>(2, *gc, *gf, [&](HostAxis<T> *h) { return build_deinterp<T>(*gc, *gf, h); }, st, &d)) return 1;
  return run_gather<T, 0>(dst, gc->n, src, gf->n, *d, (T) scale, "deinterp_add", st);
}

// ---- windows -------------------------------------------------------------------------------------------------------
// Window<T>::contains on the cell centroids (window.hpp), per axis: [lo, hi) or its wrapped complement form. The centroid
// is formed as on the host (coord * cellSize, + offsetLower, + cellSize / 2: three separately rounded operations).
template <typename T> struct WindowGeo {
  T dx, half_dx, lower[3], lo[3], hi[3];
};

template <typename T> __global__ void zero_window_kernel(T *f, int n, WindowGeo<T> w, int inside) {
  auto in_axis = [&](int a, int c) {
    const T x = add_rn(add_rn(mul_rn((T) c, w.dx), w.lower[a]), w.half_dx);
    return (w.hi[a] >= w.lo[a]) ? (x < w.hi[a] && x >= w.lo[a]) : (x < w.hi[a] || x >= w.lo[a]);
  };
  const long nrows = (long) n * n;
  for (long r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int x = (int) (r / n), y = (int) (r - (long) x * n);
    const bool xy = in_axis(0, x) && in_axis(1, y);
    T *row = f + r * (n + 2);
    for (int z = threadIdx.x; z < n; z += blockDim.x) {
      const int contained = (xy && in_axis(2, z)) ? 1 : 0;
      if (contained == inside) row[z] = (T) 0;
    }
  }
}

template <typename T>
static int zero_window_t(void *field, const gx_grid_desc *g, const gx_grid_desc *win, int inside, cudaStream_t st) {
  // Grid::getWindow (grid.hpp:502-504) -> Window(L, lower, upper): wrapCorners, then contains() on cell centroids
  const int n = g->n;
  WindowGeo<T> w;
  w.dx = (T) g->cell_size; w.half_dx = (T) g->cell_size / 2;
  const T L = (T) win->domain_size;
  auto wrap = [L](T v) { if (v < 0) v += L; if (v >= L) v -= L; return v; };
  for (int a = 0; a < 3; ++a) {
    w.lower[a] = (T) g->offset_lower[a];
    const T thisGridSize = (T) win->cell_size * (T) win->n;
    w.lo[a] = wrap((T) win->offset_lower[a]);
    const T hi = (T) win->offset_lower[a] + thisGridSize;
    w.hi[a] = std::nextafter(wrap(std::nextafter(hi, -std::numeric_limits<T>::infinity())), std::numeric_limits<T>::infinity());
  }
  const long nrows = (long) n * n;
  zero_window_kernel<T><<<(int) (nrows < 148 * 16 ? nrows : 148 * 16), 256, 0, st>>>((T *) field, n, w, inside ? 1 : 0);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

// ---- flagged-cell masks -----------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> __global__ void __launch_bounds__(256)
masked_sum_kernel(const T *__restrict__ f, const unsigned char *__restrict__ mask, int n, double *partials) {
  const long nrows = (long) n * n;
  double acc = 0;
  for (long r = blockIdx.x; r < nrows; r += gridDim.x) {
    const T *row = f + r * (n + 2);
    const unsigned char *m = mask + r * n;
    for (int z = threadIdx.x; z < n; z += blockDim.x) if (m[z]) acc += (double) row[z];
  }
  __shared__ double s[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? s[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

__global__ void partial_sum_kernel(const double *partials, int n, double *out) {
  __shared__ double s[32];
  double v = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) *out = v;
  }
}

// x <- mask ? (x - shift) * scale : 0   (windowOperator + the mean/size step of varianceOperator)
template <typename T> __global__ void masked_shift_scale_kernel(T *f, const unsigned char *mask, int n, T shift, T scale) {
  const long nrows = (long) n * n;
  for (long r = blockIdx.x; r < nrows; r += gridDim.x) {
    T *row = f + r * (n + 2);
    const unsigned char *m = mask + r * n;
    for (int z = threadIdx.x; z < n; z += blockDim.x) row[z] = m[z] ? (row[z] - shift) * scale : (T) 0;
  }
}

template <typename T> static int masked_sum_t(const void *f, const unsigned char *mask, int n, double *host_out, cudaStream_t st) {
  const long nrows = (long) n * n;
  const int grid = (int) (nrows < 148 * 8 ? nrows : 148 * 8);
  static double *part = nullptr;  // 148 * 8 partials + the result; the call returns only after its result has been read
  static int part_device = -1;
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  if (part_device != dev) { GX_CUDA_CHECK(cudaMalloc(&part, (148 * 8 + 1) * sizeof(double))); part_device = dev; }
  masked_sum_kernel<T><<<grid, 256, 0, st>>>((const T *) f, mask, n, part);
  partial_sum_kernel<<<1, 1024, 0, st>>>(part, grid, part + 148 * 8);
  count_launch(2);
  GX_LAUNCH_CHECK();
  GX_CUDA_CHECK(cudaMemcpyAsync(host_out, part + 148 * 8, sizeof(double), cudaMemcpyDeviceToHost, st));
  GX_CUDA_CHECK(cudaStreamSynchronize(st));
  return 0;
}

template <typename T>
static int masked_shift_scale_t(void *f, const unsigned char *mask, int n, double shift, double scale, cudaStream_t st) {
  const long nrows = (long) n * n;
  masked_shift_scale_kernel<T><<<(int) (nrows < 148 * 16 ? nrows : 148 * 16), 256, 0, st>>>((T *) f, mask, n, (T) shift, (T) scale);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

// ---- F <- scale * F / k^2 on the independent slots, 0 at k = 0 (VelocityModification::computeVelocity real-space-derivative
// branch, linearmodification.hpp:242-262; cosmology::densityToPotential) -------------------------------------------------
template <typename T> struct ScaleInvK2F {
  cplx<T> *f; int n; T kmin, scale; int potential_form;
  __device__ __forceinline__ void init() const {}
  __device__ __forceinline__ void operator()(int ix, int iy, int iz, size_t s) const {
    if (!slot_is_independent(n, ix, iy, iz)) return;
    T kx, ky, kz, k2;
    slot_k<T>(n, ix, iy, iz, kmin, kx, ky, kz, k2);
    cplx<T> v = f[s];
    if (potential_form) {
      // densityToPotential (parameters.hpp:208-225): ksq_inv = 1/ksq, set to 0 where the SECOND wavenumber component is
      // zero (the reference tests its lambda argument `k2`, i.e. ky, not ksq), result = scale * val * ksq_inv
      const T inv = (ky == (T) 0) ? (T) 0 : div_rn((T) 1, k2);
      f[s] = mk<T>(mul_rn(mul_rn(scale, v.x), inv), mul_rn(mul_rn(scale, v.y), inv));
    } else {
      f[s] = (k2 == (T) 0) ? mk<T>(0, 0) : mk<T>(div_rn(mul_rn(scale, v.x), k2), div_rn(mul_rn(scale, v.y), k2));
    }
  }
};

template <typename T> static int scale_inv_k2_t(void *field, int n, double L, double scale, int potential_form, cudaStream_t st) {
  ScaleInvK2F<T> f;
  f.f = (cplx<T> *) field; f.n = n; f.kmin = (T) (2. * M_PI / (double) (T) L); f.scale = (T) scale; f.potential_form = potential_form;
  return launch_slot_sweep(n, f, st, 0, "scale_inv_k2");
}

}  // namespace gx

using namespace gx;

#define GX_DISPATCH(dtype, CALL)                              \
  do {                                                        \
    if ((dtype) == GX_F64) { using T = double; return CALL; } \
    if ((dtype) == GX_F32) { using T = float; return CALL; }  \
    return gx::fail("bad dtype");                             \
  } while (0)

extern "C" {

int gx_interp_add(void *dst, const gx_grid_desc *dst_grid, const void *src, const gx_grid_desc *src_grid, int dtype,
                  void *stream) {
  if (check_desc(dst_grid, "gx_interp_add") || check_desc(src_grid, "gx_interp_add")) return 1;
  if (dst == src) return gx::fail("gx_interp_add: source and destination alias");
  GX_DISPATCH(dtype, interp_add_t<T>(dst, dst_grid, src, src_grid, (cudaStream_t) stream));
}

int gx_deinterp_add(void *dst_coarse, const gx_grid_desc *coarse_grid, const void *src_fine, const gx_grid_desc *fine_grid,
                    int dtype, double scale, void *stream) {
  if (check_desc(coarse_grid, "gx_deinterp_add") || check_desc(fine_grid, "gx_deinterp_add")) return 1;
  if (!(fine_grid->cell_size < coarse_grid->cell_size)) return gx::fail("gx_deinterp_add: the source must be the finer grid");
  GX_DISPATCH(dtype, deinterp_add_t<T>(dst_coarse, coarse_grid, src_fine, fine_grid, scale, (cudaStream_t) stream));
}

int gx_zero_window(void *field, const gx_grid_desc *grid, int dtype, const gx_grid_desc *window_grid, int inside,
                   void *stream) {
  if (check_desc(grid, "gx_zero_window") || check_desc(window_grid, "gx_zero_window")) return 1;
  GX_DISPATCH(dtype, zero_window_t<T>(field, grid, window_grid, inside, (cudaStream_t) stream));
}

int gx_masked_sum(const void *field, const unsigned char *mask_n3, int n, int dtype, double *host_out, void *stream) {
  GX_DISPATCH(dtype, masked_sum_t<T>(field, mask_n3, n, host_out, (cudaStream_t) stream));
}

int gx_masked_shift_scale(void *field, const unsigned char *mask_n3, int n, int dtype, double shift, double scale,
                          void *stream) {
  GX_DISPATCH(dtype, masked_shift_scale_t<T>(field, mask_n3, n, shift, scale, (cudaStream_t) stream));
}

int gx_scale_inv_k2(void *field, int n, int dtype, double grid_length, double scale, int potential_form, void *stream) {
  if (n < 4 || n > 4096 || (n & 1)) return gx::fail("grid size must be even and in [4,4096]");
  GX_DISPATCH(dtype, scale_inv_k2_t<T>(field, n, grid_length, scale, potential_form, (cudaStream_t) stream));
}

}  // extern "C"

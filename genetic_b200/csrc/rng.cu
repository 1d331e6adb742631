// White noise on the device, draw-for-draw identical to the reference's GSL stream.
//
// Replaces fields::RandomFieldGenerator::draw (randomfieldgenerator.hpp:122-139):
//   drawRandomForSpecifiedGrid        :181-197  N^3 gaussians in cell order, then r2c and Re(DC) = 0
//   drawRandomForSpecifiedGridFourier :203-260  square k-shells, serial (one stream) or per-shell streams
//   drawOneFourierMode                :144-171  two consecutive gaussians per mode, the "also write -k" quirk,
//                                               later writes overwrite earlier ones
//
// Parallelisation (rng_core.cuh has the arithmetic):
//   1. every stream (one per shell in parallel mode, one in total otherwise) is cut into chunks of kBc
//      increments (kBc*24 raw outputs); a thread jumps to its chunk with the LCG skip-ahead;
//   2. pass A: the thread generates its chunk and computes the chunk's *transfer function*: for each of the
//      four possible positions at which the first ziggurat attempt of the chunk can start, where the parse
//      leaves the chunk and how many gaussians it accepted (attempts consume 2-4 outputs; the four parses
//      merge after a few attempts, so this costs ~1.1 parses);
//   3. a 64-ary tree scan composes the transfer functions (associative) and hands every chunk its true entry
//      position and the index of its first gaussian;
//   4. pass B regenerates the chunk, parses it once and delivers gaussian g of the stream straight to its
//      storage slot -- for the Fourier orders as a gather-free scatter in which a draw is written only if it
//      is the last writer of its slot (a mode and its conjugate partner can both be drawn in a shell);
//   5. accepted tail draws (~6e-4 of all) depend on log(): their raw integers are listed and the values are
//      recomputed on the HOST with the C library's log() -- exactly the call GSL makes -- and patched in, so
//      the field is bit-identical to what the reference produces on the same machine.
#include "../../include/genetic_b200.h"
#include "rng_core.cuh"
#include "rng_staged_core.cuh"
#include "zig_tables.h"  // generated constant tables (GSL's published 128-level ziggurat)
#include "zeldovich.cuh"   // build_transfer_table
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstdlib>
#include <thread>

namespace gx {
void release_spectral_workspace();   // spectral.cu: the |k|^2 table
void release_reduce_workspace();     // reduce.cu: partial sums
void pool_release();                 // mempool.cu: idle field buffers
namespace fftgen { void release_workspace(); }                                              // fft_generic.cu: twiddle tables
namespace fft { template <typename T> int fft3d_any(T *field, int n, int dir, cudaStream_t st); }
void release_multigrid_workspace();  // multigrid.cu: cached cross-grid tables
template <typename T> int mirror(T *field, int n, cudaStream_t st);
namespace fft { template <typename T> int fft3d(T *data, int n, int dir, cudaStream_t st); }

namespace rng {

constexpr int kThreads = 128;
constexpr int kJumpDigits = 3, kJumpRadix = 512;   // chunk index -> jump: three base-512 digits, one table product each

__device__ unsigned int d_zig_ktab[128];
__device__ double d_zig_wtab[128];
__device__ double d_zig_ytab[128];

struct TailEvent {
  long long g;       // stream-local gaussian index
  int stream;
  unsigned int u1;   // raw integer behind U1
  int positive;      // sign of the draw
};

enum { MODE_REAL = 0, MODE_FOURIER_SERIAL = 1, MODE_FOURIER_PARALLEL = 2, MODE_LIST = 3 };

template <typename T> struct Dest {
  int mode;
  int n;
  int reverse;
  double sigma;      // FloatType sigma = 1.0 / sqrt(2.0) (randomfieldgenerator.hpp:213), widened to double
  T *field;          // padded field (real or Fourier layout), or this rank's slab of it
  double *list;      // MODE_LIST: plain gaussian list (test hook)
  // slab ownership (multi-GPU): Fourier orders keep the modes with storage index iy in [iy0, iy0 + ny) in a
  // [n][ny][n/2+1] array; the real order keeps rows (x*n + y) in [row0, row0 + nrows). Whole field: 0, n / 0, n*n.
  int iy0, ny;
  long long row0, nrows;
  // optional factor per mode, table[k1^2 + k2^2 + k3^2] (sqrt of the covariance [/ k^2]): Field::applyTransferFunction
  // folded into the delivery -- the same rounded product per component as the separate sweep, one field sweep less
  const T *wtab;
};

// ---- delivery of gaussian g of a stream to its destination ---------------------------------------------------
// calls before shell ks in the serial order: sum_{j<ks} 3 (2j)^2 = 2 (ks-1) ks (2ks-1)
__device__ __forceinline__ long long calls_before_shell(long long ks) { return 2 * (ks - 1) * ks * (2 * ks - 1); }

// index of the drawOneFourierMode call that draws mode (k1,k2,k3) inside its shell, or -1 if it is never drawn
__device__ __forceinline__ long long call_index_in_shell(int ks, int a, int b, int c) {
  // drawn iff exactly one coordinate equals +ks and the other two lie in [-ks, ks-1]
  const int np = (a == ks) + (b == ks) + (c == ks);
  if (np != 1) return -1;
  int face, p, q;
  if (a == ks) { face = 0; p = b; q = c; }
  else if (b == ks) { face = 1; p = a; q = c; }
  else { face = 2; p = a; q = b; }
  if (p < -ks || p >= ks || q < -ks || q >= ks) return -1;
  return ((long long) (p + ks) * (2 * ks) + (q + ks)) * 3 + face;
}

template <typename T>
__device__ __forceinline__ void store_part(T *field, int n, int kx, int ky, int kz, int part, T v) {
  // kz >= 0 here
  if (ky < 0) ky += n;
  if (kx < 0) kx += n;
  const size_t slot = (size_t) kz + (size_t) (n / 2 + 1) * ((size_t) ky + (size_t) n * kx);
  field[2 * slot + part] = v;
}

// Position of a gaussian in its draw order. Initialised once per chunk from the gaussian index (divisions), then
// advanced incrementally: the draw loops of the reference are plain nested counters.
struct Cursor {
  int ks, p, q, face, comp;  // Fourier orders: shell, (k1, k2) loop counters, which of the 3 modes, re/im draw
  long long row;             // real order: row = x*n + y, z
  int z;
};

template <typename T>
__device__ __forceinline__ void cursor_init(const Dest<T> &d, const StreamDesc &sd, long long g, Cursor &cu) {
  cu.ks = cu.p = cu.q = cu.face = cu.comp = 0; cu.row = 0; cu.z = 0;
  if (d.mode == MODE_LIST) return;
  if (d.mode == MODE_REAL) {
    cu.row = g / d.n;
    cu.z = (int) (g - cu.row * d.n);
    return;
  }
  long long c = g >> 1;
  cu.comp = (int) (g & 1);
  int ks;
  if (d.mode == MODE_FOURIER_PARALLEL) ks = sd.shell;
  else {
    ks = (int) cbrt((double) c / 4.0) + 1;
    while (ks > 1 && calls_before_shell(ks) > c) --ks;
    while (calls_before_shell(ks + 1) <= c) ++ks;
    c -= calls_before_shell(ks);
  }
  cu.ks = ks;
  cu.face = (int) (c % 3);
  const long long t = c / 3;
  cu.p = (int) (t / (2 * ks)) - ks;
  cu.q = (int) (t % (2 * ks)) - ks;
}

template <typename T> __device__ __forceinline__ void cursor_advance(const Dest<T> &d, Cursor &cu) {
  if (d.mode == MODE_LIST) return;
  if (d.mode == MODE_REAL) {
    const bool wrap = (cu.z + 1 == d.n);
    cu.z = wrap ? 0 : cu.z + 1;
    cu.row += wrap ? 1 : 0;
    return;
  }
  // nested counters (comp fastest, then face, q, p, shell) as carries instead of branches
  const int c0 = cu.comp;                       // carry out of comp
  cu.comp ^= 1;
  const int face1 = cu.face + c0;
  const int c1 = (face1 == 3);
  cu.face = c1 ? 0 : face1;
  const int q1 = cu.q + c1;
  const int c2 = (q1 == cu.ks);
  cu.q = c2 ? -cu.ks : q1;
  const int p1 = cu.p + c2;
  const int c3 = (p1 == cu.ks);
  // serial order: on to the next shell (a per-shell stream simply ends here)
  cu.ks += c3;
  cu.p = c3 ? -cu.ks : p1;
  cu.q = c3 ? -cu.ks : cu.q;
}

// Where the mode drawn at cursor `cu` is stored (Fourier orders). Branch-free on purpose: neighbouring lanes sit at
// different (face, re/im, sign of kz) phases of the draw loops, so every branch here used to be executed serially for
// 6+ lane groups (ncu: 9 of 32 threads active on average).
struct ModeSlot {
  size_t slot, slot2;   // complex slot of the kz >= 0 representative; slot of -k on the kz = 0 plane
  bool write, write2;   // last-writer-wins and slab ownership already applied
  bool neg;             // the representative is the conjugate (k3 < 0)
  int m;                // k1^2 + k2^2 + k3^2
};

template <typename T> __device__ __forceinline__ ModeSlot mode_slot(const Dest<T> &d, const Cursor &cu) {
  const int n = d.n;
  const int ks = cu.ks, p = cu.p, q = cu.q, face = cu.face;
  // drawOneFourierMode is called for (ks,p,q), (p,ks,q), (p,q,ks) in turn
  const int k1 = (face == 0) ? ks : p;
  const int k2 = (face == 0) ? p : ((face == 1) ? ks : q);
  const int k3 = (face == 2) ? ks : q;
  // last-writer-wins: the conjugate partner -k is drawn in the same shell iff exactly one of its coordinates equals
  // +ks (i.e. one of ours equals -ks) and the other two lie in [-ks, ks-1] (ours in [-ks+1, ks])
  const int a = -k1, b = -k2, c3 = -k3;
  const int np = (a == ks) + (b == ks) + (c3 == ks);
  const int pface = (a == ks) ? 0 : ((b == ks) ? 1 : 2);
  const int pp = (pface == 0) ? b : a;
  const int pq = (pface == 2) ? b : c3;
  const bool partner_drawn = (np == 1) && (pp >= -ks) && (pp < ks) && (pq >= -ks) && (pq < ks);
  const int twoks = 2 * ks;
  const int c_self = ((p + ks) * twoks + (q + ks)) * 3 + face;      // < 3 * 4096^2: fits an int
  const int c_partner = ((pp + ks) * twoks + (pq + ks)) * 3 + pface;
  const bool write = !(partner_drawn && c_partner > c_self);
  // setFourierCoefficient(k, val): stored at the kz >= 0 representative, imaginary part negated if conjugated;
  // the explicit second write, setFourierCoefficient(-k, conj(val)), lands on a different slot only when k3 == 0
  ModeSlot m;
  m.m = k1 * k1 + k2 * k2 + k3 * k3;
  m.neg = k3 < 0;
  const int s1 = m.neg ? -k1 : k1, s2 = m.neg ? -k2 : k2, s3 = m.neg ? -k3 : k3;
  const int nc = n / 2 + 1;
  const int j2 = (s2 < 0 ? s2 + n : s2) - d.iy0;
  m.slot = (size_t) s3 + (size_t) nc * ((size_t) j2 + (size_t) d.ny * (size_t) (s1 < 0 ? s1 + n : s1));
  m.write = write && j2 >= 0 && j2 < d.ny;
  const int t1 = -k1, t2 = -k2;
  const int u2 = (t2 < 0 ? t2 + n : t2) - d.iy0;
  m.slot2 = (size_t) nc * ((size_t) u2 + (size_t) d.ny * (size_t) (t1 < 0 ? t1 + n : t1));
  m.write2 = write && k3 == 0 && u2 >= 0 && u2 < d.ny;
  return m;
}

template <typename T>
__device__ __forceinline__ void deliver_at(const Dest<T> &d, const Cursor &cu, long long g, double value) {
  if (d.mode == MODE_LIST) { d.list[g] = value; return; }
  const int n = d.n;
  if (d.mode == MODE_REAL) {
    const long long r = cu.row - d.row0;
    if (r >= 0 && r < d.nrows) d.field[r * (n + 2) + cu.z] = (T) value;
    return;
  }
  const ModeSlot m = mode_slot<T>(d, cu);
  // value of this component in the field's type: norm * gaussian with norm = FloatType(1/sqrt(2))
  T v = (T) GX_DMUL(d.sigma, value);
  if (d.wtab) v = (T) GX_DMUL(v, __ldg(d.wtab + m.m));
  // normal order: (a, b) -> (re, im); reverse: (b, a)
  const int part = d.reverse ? (1 - cu.comp) : cu.comp;
  const T vconj = (part == 1) ? -v : v;
  if (m.write) d.field[2 * m.slot + part] = m.neg ? vconj : v;
  if (m.write2) d.field[2 * m.slot2 + part] = vconj;
}

// Both draws of one drawOneFourierMode call at once (the cursor is the same for the two): one 2-component store per
// slot instead of two scattered scalar ones, and the index arithmetic once per mode.
template <typename T>
__device__ __forceinline__ void deliver_pair(const Dest<T> &d, const Cursor &cu, double first, double second) {
  const ModeSlot m = mode_slot<T>(d, cu);
  T a = (T) GX_DMUL(d.sigma, first), b = (T) GX_DMUL(d.sigma, second);
  if (d.wtab) { const T w = __ldg(d.wtab + m.m); a = (T) GX_DMUL(a, w); b = (T) GX_DMUL(b, w); }
  const T re = d.reverse ? b : a, im = d.reverse ? a : b;
  cplx<T> *f = reinterpret_cast<cplx<T> *>(d.field);
  if (m.write) f[m.slot] = mk<T>(re, m.neg ? -im : im);
  if (m.write2) f[m.slot2] = mk<T>(re, -im);
}

template <typename T>
__device__ void deliver(const Dest<T> &d, const StreamDesc &sd, long long g, double value) {
  Cursor cu;
  cursor_init<T>(d, sd, g, cu);
  deliver_at<T>(d, cu, g, value);
}

// ---- chunk kernels -------------------------------------------------------------------------------------------
struct ChunkParams {
  const StreamDesc *streams;
  int n_streams;
  int n_chunks;
  const Big *jump;            // jump[d * 512 + i] = a^(397 * kBc * i * 512^d)
  ChunkFn *fn;                // pass A output
  const ScanState *scan;      // pass B input
  unsigned long long *stream_end;  // pass B: raw position after the last wanted gaussian (per stream)
  TailEvent *tails;
  unsigned int *tail_count;
  unsigned int tail_capacity;
  const int *list;            // pass B on a slab: the chunks that can touch this rank's slots (nullptr: all chunks)
  int n_list;
  int first, count;           // list == nullptr: the kernel covers chunks [first, first + count)
  ChunkFn *fn_out;            // pass A writes fn_out[chunk - first]
};

// jump[d][i] = J^(i * 512^d), J = a^(397 * kBc): at most three 576-bit products per chunk start (the bitwise table of
// round 1 cost ~7.5 at 2048^3, a quarter of the chunk's generator work)
__device__ __forceinline__ State chunk_start_state(const StreamDesc &sd, int q, const Big *jump) {
  if (sd.b0 == 0 && q == 0) return sd.s0;
  Big X = sd.xs;
#pragma unroll 1
  for (int d = 0; d < kJumpDigits; ++d) {
    const int digit = (q >> (9 * d)) & (kJumpRadix - 1);
    if (digit) X = modmul(X, jump[d * kJumpRadix + digit]);
  }
  return big_to_state(X);
}

template <int PASS, typename T, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) rng_chunk_kernel(ChunkParams P, Dest<T> dest) {
  __shared__ unsigned int s_k[128];
  __shared__ double s_w[128], s_y[128];
  __shared__ uint32_t ring[64][kThreads];  // raw output p of the chunk lives in slot p & 63 (the parse looks back < 28 outputs)
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { s_k[i] = d_zig_ktab[i]; s_w[i] = d_zig_wtab[i]; s_y[i] = d_zig_ytab[i]; }
  __syncthreads();
  // No lane leaves before the end: every increment starts with a __syncwarp so that the 397-step generator advance
  // runs with the whole warp converged (lanes that had left early or sat in different parse branches used to run it
  // in 5-lane groups -- 70% of all issued instructions of pass B, ncu r1p).
  const int slot_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = slot_raw < (P.list ? P.n_list : P.count);
  const int chunk = in_range ? (P.list ? P.list[slot_raw] : P.first + slot_raw) : P.n_chunks - 1;
  bool live = in_range;
  int sid = 0;
  {  // the stream this chunk belongs to: last stream with first_chunk <= chunk
    int lo = 0, hi = P.n_streams - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (P.streams[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
    }
    sid = lo;
  }
  const StreamDesc &sd = P.streams[sid];
  const int q = chunk - sd.first_chunk;
  const bool first = (q == 0);
  const int tx = threadIdx.x;

  int pos[4];
  long long cnt[4];
  int rep[4];          // representative cursor after a merge (-1: own)
  long long delta[4];  // cnt[e] - cnt[rep[e]] at merge time
  int ncur;
  long long g = 0;
  if (PASS == 0) {
    if (first) { ncur = 1; pos[0] = sd.entry0; }
    else { ncur = 4; for (int e = 0; e < 4; ++e) pos[e] = e; }
    for (int e = 0; e < 4; ++e) { cnt[e] = 0; rep[e] = -1; delta[e] = 0; }
  } else {
    const ScanState ss = P.scan[chunk];
    ncur = 1;
    pos[0] = first ? sd.entry0 : ss.entry;
    g = first ? 0 : ss.base;
    if (g >= sd.n_gauss) live = false;  // nothing wanted from this chunk
  }

  Cursor cu;
  if (PASS == 1) cursor_init<T>(dest, sd, g, cu);
  const bool pairs = (PASS == 1) && (dest.mode == MODE_FOURIER_SERIAL || dest.mode == MODE_FOURIER_PARALLEL);
  bool have_held = false;  // the first draw of a mode waits for the second (same thread unless the chunk ends in between)
  double held = 0.0;

  State st = chunk_start_state(sd, q, P.jump);

  for (int bi = 0; bi <= kBc; ++bi) {
    __syncwarp();
    advance_block(st);
    const int rbase = 24 * bi;
#pragma unroll
    for (int j = 0; j < 24; ++j) ring[(rbase + j) & 63][tx] = block_output(st, j);
    const int avail = 24 * (bi + 1);
    const int limit = min(avail - 3, kChunkLen);
    for (int e = 0; e < ncur; ++e) {
      if (PASS == 0 && rep[e] >= 0) continue;
      int p = pos[e];
      while (live && p < limit) {
        const Attempt a = zig_attempt(ring[p & 63][tx], ring[(p + 1) & 63][tx], ring[(p + 2) & 63][tx],
                                      ring[(p + 3) & 63][tx], s_k, s_w, s_y);
        p += a.consumed;
        if (a.accepted) {
          if (PASS == 0) cnt[e]++;
          else {
            if (g < sd.n_gauss) {
              if (!pairs) deliver_at<T>(dest, cu, g, a.value);
              else if (cu.comp == 0) { held = a.value; have_held = true; }
              else if (have_held) { deliver_pair<T>(dest, cu, held, a.value); have_held = false; }
              else deliver_at<T>(dest, cu, g, a.value);  // its partner was the last draw of the previous chunk
              if (a.tail) {
                const unsigned int slot = atomicAdd(P.tail_count, 1u);
                if (slot < P.tail_capacity) {
                  TailEvent ev; ev.g = g; ev.stream = sid; ev.u1 = a.tail_u1; ev.positive = a.value > 0;
                  P.tails[slot] = ev;
                }
              }
              if (g == sd.n_gauss - 1)
                P.stream_end[sid] = (unsigned long long) (sd.b0 + (long long) q * kBc) * 24ull + (unsigned long long) p;
            }
            ++g;
            cursor_advance<T>(dest, cu);
          }
        }
      }
      pos[e] = p;
    }
    if (PASS == 0 && ncur == 4) {
      // merge cursors that have reached the same position
      for (int e = 1; e < 4; ++e) {
        if (rep[e] >= 0) continue;
        for (int r = 0; r < e; ++r) {
          if (rep[r] < 0 && pos[r] == pos[e]) { rep[e] = r; delta[e] = cnt[e] - cnt[r]; break; }
        }
      }
    }
    if (PASS == 1 && g >= sd.n_gauss) live = false;
  }

  if (PASS == 1 && have_held) {  // the chunk ended between the two draws of a mode: the cursor already points at comp 1
    Cursor c0 = cu;
    c0.comp = 0;
    deliver_at<T>(dest, c0, g - 1, held);
  }

  if (PASS == 0 && in_range) {
    ChunkFn f;
    f.reset = first ? 1 : 0;
    for (int e = 0; e < 4; ++e) {
      int r = first ? 0 : e;
      long long dsum = 0;
      while (rep[r] >= 0) { dsum += delta[r]; r = rep[r]; }  // merges can chain (3 -> 1 -> 0)
      f.exit[e] = (unsigned char) (pos[r] - kChunkLen);
      f.count[e] = cnt[r] + dsum;
    }
    P.fn_out[chunk - P.first] = f;
  }
}

// ---- slab ownership: which chunks of the per-shell streams can write into ky storage rows [iy0, iy0 + ny) ---------------
// A draw of shell ks is stored at ky = +-k2 with k2 = p (mode (ks,p,q)), ks (mode (p,ks,q)) or q (mode (p,q,ks)); a chunk
// holds at most kChunkLen / 2 gaussians, i.e. a short run of the (p, q) loop. The test is conservative (intervals, both
// signs); chunks it rejects are not generated at all in pass B. The appended order is irrelevant: every chunk is
// delivered independently and overwrites are resolved analytically (last writer wins by call index).
__global__ void need_kernel(const StreamDesc *streams, int n_streams, int n_chunks, const ScanState *scan, int n, int iy0,
                            int ny, int *list, unsigned int *count) {
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= n_chunks) return;
  int lo = 0, hi = n_streams - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (streams[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
  }
  const StreamDesc &sd = streams[lo];
  const long long g0 = (chunk == sd.first_chunk) ? 0 : scan[chunk].base;
  if (g0 >= sd.n_gauss) return;
  long long g1 = g0 + kChunkLen / 2 + 3;   // past the last gaussian the chunk can hold (>= 2 outputs each, entry offset <= 3)
  bool need = false;
  if (g1 >= sd.n_gauss) { g1 = sd.n_gauss; need = true; }  // the chunk that records the end of the stream always runs
  // signed ky values of the slab: storage index j <= n/2 is ky = j, above it ky = j - n (ky = n/2 is never drawn)
  const int s0 = (iy0 + ny <= n / 2) ? iy0 : iy0 - n, s1 = s0 + ny - 1;
  auto owned = [&](int a, int b) { return a <= s1 && b >= s0; };  // does [a, b] meet the slab?
  const int ks = sd.shell, two = 2 * ks;
  const long long t0 = (g0 / 2) / 3, t1 = ((g1 - 1) / 2) / 3;
  const int pa = (int) (t0 / two) - ks, pb = (int) (t1 / two) - ks;
  int qa = (int) (t0 % two) - ks, qb = (int) (t1 % two) - ks;
  if (pa != pb) { qa = -ks; qb = ks - 1; }  // the run wraps into the next p: every q can occur
  // modes (ks,p,q) and (p,ks,q) have kz = q: stored at ky = k2 for q >= 0, at the conjugate ky = -k2 for q < 0, and at
  // both for q == 0 (the explicit second write of drawOneFourierMode); mode (p,q,ks) has kz = ks > 0: stored at ky = q.
  const bool plus = qb >= 0, minus = qa <= 0;
  need = need || (plus && (owned(pa, pb) || owned(ks, ks))) || (minus && (owned(-pb, -pa) || owned(-ks, -ks))) ||
         owned(qa, qb);
  if (need) list[atomicAdd(count, 1u)] = chunk;
}

// residency of the chunk kernels: 5 CTAs per SM (96 registers) by default -- measured at 2048^3: 4 CTAs (128 registers)
// 99 + 180 ms, 5 CTAs 96 + 168 ms, 6 CTAs (80 registers, more spills) 100 + 196 ms (profiles/r1s17_*). GX_RNG_MINB=4|6
// selects the others.
template <int PASS, typename T>
static void launch_chunk_kernel(int grid, cudaStream_t cs, const ChunkParams &P, const Dest<T> &dest) {
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("GX_RNG_MINB"); minb = e ? atoi(e) : 5; }
  if (minb == 4) rng_chunk_kernel<PASS, T, 4><<<grid, kThreads, 0, cs>>>(P, dest);
  else if (minb == 6) rng_chunk_kernel<PASS, T, 6><<<grid, kThreads, 0, cs>>>(P, dest);
  else rng_chunk_kernel<PASS, T, 5><<<grid, kThreads, 0, cs>>>(P, dest);
}

// ---- scan over chunk functions -----------------------------------------------------------------------------------
__global__ void scan_up_kernel(const ChunkFn *in, int n_in, ChunkFn *out, int n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  ChunkFn h = in[(size_t) 64 * i];
  for (int k = 1; k < 64; ++k) {
    const size_t idx = (size_t) 64 * i + k;
    if (idx >= (size_t) n_in) break;
    h = compose(h, in[idx]);
  }
  out[i] = h;
}

__global__ void scan_top_kernel(const ChunkFn *fn, int n, ScanState *st) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  ScanState s; s.base = 0; s.entry = 0;
  for (int i = 0; i < n; ++i) { st[i] = s; s = apply_fn(fn[i], s); }
}

__global__ void scan_down_kernel(const ChunkFn *child_fn, int n_child, const ScanState *parent, int n_parent,
                                 ScanState *child) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_parent) return;
  ScanState s = parent[i];
  for (int k = 0; k < 64; ++k) {
    const size_t idx = (size_t) 64 * i + k;
    if (idx >= (size_t) n_child) break;
    child[idx] = s;
    s = apply_fn(child_fn[idx], s);
  }
}

// ---- patch kernel: re-deliver tail draws with host-computed values -------------------------------------------------
template <typename T>
__global__ void tail_patch_kernel(const TailEvent *ev, const double *values, unsigned int n, const StreamDesc *streams,
                                  Dest<T> dest) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  deliver<T>(dest, streams[ev[i].stream], ev[i].g, values[i]);
}

// test hook: plain raw outputs of whole chunks
__global__ void raw_chunks_kernel(StreamDesc sd, const Big *jump, long long q0, int nch, uint32_t *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nch) return;
  State st = chunk_start_state(sd, (int) (q0 + i), jump);
  for (int bi = 0; bi < kBc; ++bi) {
    advance_block(st);
    for (int j = 0; j < 24; ++j) out[((size_t) i * kBc + bi) * 24 + j] = block_output(st, j);
  }
}

template <typename T> __global__ void zero_dc_real_kernel(T *field) { field[0] = (T) 0; }

// accepted tail draws: x = R - log(U1) / R with the C library's log() -- the call gsl_ran_gaussian_ziggurat makes. ~6e-4 of
// all draws (5 million at 2048^3): spread over the host threads.
struct TailEvent;
static void host_tail_values(const TailEvent *ev, double *val, unsigned int ntail);
// ---- host driver ------------------------------------------------------------------------------------------------------
static Big *g_jump = nullptr;
static int g_jump_device = -1;
static int ensure_jump_table() {
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_jump && dev == g_jump_device) return 0;
  g_jump = nullptr;  // (a table on another device stays allocated there until that device is reset)
  g_jump_device = dev;
  std::vector<Big> h((size_t) kJumpDigits * kJumpRadix);
  Big base = modpow(big_a(), 397ull * kBc);
  for (int d = 0; d < kJumpDigits; ++d) {
    Big acc = big_one();
    for (int i = 0; i < kJumpRadix; ++i) { h[(size_t) d * kJumpRadix + i] = acc; acc = modmul(acc, base); }
    base = acc;  // base^512
  }
  GX_CUDA_CHECK(cudaMemcpyToSymbol(d_zig_ktab, gx_zig_ktab, sizeof(gx_zig_ktab)));
  GX_CUDA_CHECK(cudaMemcpyToSymbol(d_zig_wtab, gx_zig_wtab, sizeof(gx_zig_wtab)));
  GX_CUDA_CHECK(cudaMemcpyToSymbol(d_zig_ytab, gx_zig_ytab, sizeof(gx_zig_ytab)));
  GX_CUDA_CHECK(cudaMalloc(&g_jump, sizeof(Big) * h.size()));
  GX_CUDA_CHECK(cudaMemcpy(g_jump, h.data(), sizeof(Big) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}

static void host_tail_values(const TailEvent *ev, double *val, unsigned int ntail) {
  const double R = GX_ZIG_R;
  auto work = [&](unsigned int i0, unsigned int i1) {
    for (unsigned int i = i0; i < i1; ++i) {
      const double U1 = 1.0 - (double) ev[i].u1 / 16777216.0;
      const double x = R - std::log(U1) / R;
      val[i] = ev[i].positive ? x : -x;
    }
  };
  unsigned int nt = std::thread::hardware_concurrency();
  if (nt > 32) nt = 32;
  if (nt < 2 || ntail < 65536) { work(0, ntail); return; }
  std::vector<std::thread> pool;
  const unsigned int per = (ntail + nt - 1) / nt;
  for (unsigned int t = 0; t < nt; ++t) {
    const unsigned int i0 = t * per, i1 = std::min(ntail, i0 + per);
    if (i0 < i1) pool.emplace_back(work, i0, i1);
  }
  for (auto &th : pool) th.join();
}

static long long calls_before_shell_host(long long ks) { return 2 * (ks - 1) * ks * (2 * ks - 1); }

static long long raw_budget(long long n_gauss) {
  // measured mean 2.0528 raw outputs per gaussian (sd ~0.25 per gaussian); margin verified after the fact
  return (long long) (2.058 * (double) n_gauss + 12.0 * sqrt((double) n_gauss) + 128.0);
}

struct StreamReq { unsigned long seed; unsigned long long start_raw; long long n_gauss; int shell; };

// Grow-only device/pinned workspace: a draw needs ~50 bytes per 1536 raw outputs (chunk functions and scan
// states) plus the tail list; allocating and freeing that per call cost more than the kernels at 512^3.
struct Workspace {
  void *dev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t dev_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void *pin[3] = {nullptr, nullptr, nullptr};
  size_t pin_cap[3] = {0, 0, 0};
  int device = -1;
};
static Workspace g_ws;

static void ws_release() {
  for (int i = 0; i < 8; ++i) { if (g_ws.dev[i]) cudaFree(g_ws.dev[i]); g_ws.dev[i] = nullptr; g_ws.dev_cap[i] = 0; }
  for (int i = 0; i < 3; ++i) { if (g_ws.pin[i]) cudaFreeHost(g_ws.pin[i]); g_ws.pin[i] = nullptr; g_ws.pin_cap[i] = 0; }
}

static int ws_bind_device() {  // call before the first ws_dev / ws_pin of a draw
  int dev = 0;
  GX_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != g_ws.device) { ws_release(); g_ws.device = dev; }
  return 0;
}

static int ws_dev(int slot, size_t bytes, void **out) {
  if (g_ws.dev_cap[slot] < bytes) {
    if (g_ws.dev[slot]) GX_CUDA_CHECK(cudaFree(g_ws.dev[slot]));
    g_ws.dev[slot] = nullptr; g_ws.dev_cap[slot] = 0;
    const size_t cap = bytes + bytes / 8 + 4096;
    GX_CUDA_CHECK(cudaMalloc(&g_ws.dev[slot], cap));
    g_ws.dev_cap[slot] = cap;
  }
  *out = g_ws.dev[slot];
  return 0;
}

static int ws_pin(int slot, size_t bytes, void **out) {
  if (g_ws.pin_cap[slot] < bytes) {
    if (g_ws.pin[slot]) GX_CUDA_CHECK(cudaFreeHost(g_ws.pin[slot]));
    g_ws.pin[slot] = nullptr; g_ws.pin_cap[slot] = 0;
    const size_t cap = bytes + bytes / 8 + 4096;
    GX_CUDA_CHECK(cudaMallocHost(&g_ws.pin[slot], cap));
    g_ws.pin_cap[slot] = cap;
  }
  *out = g_ws.pin[slot];
  return 0;
}

// phase 0: everything. phase 1: pass A only, for this rank's share of the chunks, into `fn_io` (device, `fn_stride` records
// per rank). phase 2: skip pass A, take the chunk functions of all ranks from `fn_io` (all-gathered by the caller).
struct Phase { int which = 0; int world = 1, rank = 0; void *fn_io = nullptr; };

template <typename T>
static int run_streams(const std::vector<StreamReq> &reqs, Dest<T> dest, std::vector<unsigned long long> *end_raw,
                       cudaStream_t cs, Phase phase = Phase()) {
  if (ensure_jump_table()) return 1;
  const int ns = (int) reqs.size();
  if (ns == 0) return 0;
  if (ws_bind_device()) return 1;
  // stream descriptors are staged in pinned memory (slot 0) followed by the per-stream end positions
  void *pin0 = nullptr;
  if (ws_pin(0, sizeof(StreamDesc) * ns + sizeof(unsigned long long) * ns + 64, &pin0)) return 1;
  StreamDesc *sd = (StreamDesc *) pin0;
  unsigned long long *h_end = (unsigned long long *) ((char *) pin0 + sizeof(StreamDesc) * ns);
  unsigned int *h_ntail = (unsigned int *) (h_end + ns);
  const Big a = big_a();
  long long n_chunks = 0, total_gauss = 0;
  for (int s = 0; s < ns; ++s) {
    StreamDesc &d = sd[s];
    d.s0 = seed_state(reqs[s].seed);
    d.b0 = (long long) (reqs[s].start_raw / 24ull);
    d.entry0 = (int) (reqs[s].start_raw % 24ull);
    d.n_gauss = reqs[s].n_gauss;
    d.shell = reqs[s].shell;
    d.xs = state_to_big(d.s0);
    if (d.b0 > 0) d.xs = modmul(d.xs, modpow(a, 397ull * (unsigned long long) d.b0));
    d.first_chunk = (int) n_chunks;
    const long long raws = d.entry0 + raw_budget(d.n_gauss);
    n_chunks += (raws + kChunkLen - 1) / kChunkLen;
    if (n_chunks > 2000000000ll) return fail("gx_noise_draw: too many chunks");
    total_gauss += d.n_gauss;
  }
  const int n0 = (int) n_chunks;
  if (n0 == 0) return 0;
  const int per_rank = (n0 + phase.world - 1) / phase.world;  // chunk functions per rank in the all-gathered buffer

  // scan levels
  std::vector<int> level_n;
  level_n.push_back(n0);
  while (level_n.back() > 4096) level_n.push_back((level_n.back() + 63) / 64);
  const int nl = (int) level_n.size();
  size_t fn_total = 0;
  for (int l = 0; l < nl; ++l) fn_total += (size_t) level_n[l];

  const unsigned int tail_cap = (unsigned int) std::min<long long>((long long) (total_gauss * 8e-4) + 4096, 400000000ll);
  void *p_streams, *p_misc, *p_tails, *p_fn, *p_st;
  if (ws_dev(0, sizeof(StreamDesc) * ns, &p_streams)) return 1;
  if (ws_dev(1, sizeof(unsigned long long) * ns + 64, &p_misc)) return 1;
  if (ws_dev(2, sizeof(TailEvent) * (size_t) tail_cap, &p_tails)) return 1;
  if (ws_dev(3, sizeof(ChunkFn) * fn_total, &p_fn)) return 1;
  if (ws_dev(4, sizeof(ScanState) * fn_total, &p_st)) return 1;
  StreamDesc *d_streams = (StreamDesc *) p_streams;
  unsigned long long *d_end = (unsigned long long *) p_misc;
  unsigned int *d_tail_count = (unsigned int *) (d_end + ns);
  TailEvent *d_tails = (TailEvent *) p_tails;
  std::vector<ChunkFn *> d_fn(nl);
  std::vector<ScanState *> d_st(nl);
  {
    size_t off = 0;
    for (int l = 0; l < nl; ++l) { d_fn[l] = (ChunkFn *) p_fn + off; d_st[l] = (ScanState *) p_st + off; off += (size_t) level_n[l]; }
  }
  GX_CUDA_CHECK(cudaMemcpyAsync(d_streams, sd, sizeof(StreamDesc) * ns, cudaMemcpyHostToDevice, cs));
  GX_CUDA_CHECK(cudaMemsetAsync(d_end, 0xFF, sizeof(unsigned long long) * ns, cs));
  GX_CUDA_CHECK(cudaMemsetAsync(d_tail_count, 0, sizeof(unsigned int), cs));

  ChunkParams P;
  P.streams = d_streams; P.n_streams = ns; P.n_chunks = n0; P.jump = g_jump; P.fn = d_fn[0];
  P.scan = d_st[0]; P.stream_end = d_end; P.tails = d_tails; P.tail_count = d_tail_count; P.tail_capacity = tail_cap;
  P.list = nullptr; P.n_list = 0; P.first = 0; P.count = n0; P.fn_out = d_fn[0];
  const int grid = (n0 + kThreads - 1) / kThreads;
  if (phase.which == 1) {
    P.first = std::min(n0, phase.rank * per_rank);
    P.count = std::min(n0, P.first + per_rank) - P.first;
    P.fn_out = (ChunkFn *) phase.fn_io;
    if (P.count > 0) {
      GX_TIMED("rng_pass_a", cs);
      launch_chunk_kernel<0, T>((P.count + kThreads - 1) / kThreads, cs, P, dest);
      count_launch();
    }
    GX_LAUNCH_CHECK();
    return 0;
  }
  if (phase.which == 2) {
    GX_CUDA_CHECK(cudaMemcpyAsync(d_fn[0], phase.fn_io, sizeof(ChunkFn) * (size_t) n0, cudaMemcpyDeviceToDevice, cs));
  } else {
    GX_TIMED("rng_pass_a", cs);
    launch_chunk_kernel<0, T>(grid, cs, P, dest);
    count_launch();
  }
  GX_LAUNCH_CHECK();
  P.first = 0; P.count = n0;
  {
    GX_TIMED("rng_scan", cs);
    for (int l = 1; l < nl; ++l) {
      scan_up_kernel<<<(level_n[l] + 127) / 128, 128, 0, cs>>>(d_fn[l - 1], level_n[l - 1], d_fn[l], level_n[l]);
      count_launch();
    }
    scan_top_kernel<<<1, 32, 0, cs>>>(d_fn[nl - 1], level_n[nl - 1], d_st[nl - 1]);
    count_launch();
    for (int l = nl - 1; l >= 1; --l) {
      scan_down_kernel<<<(level_n[l] + 127) / 128, 128, 0, cs>>>(d_fn[l - 1], level_n[l - 1], d_st[l], level_n[l], d_st[l - 1]);
      count_launch();
    }
  }
  GX_LAUNCH_CHECK();
  int grid_b = grid;
  P.list = nullptr; P.n_list = 0;
  if (dest.mode == MODE_FOURIER_PARALLEL && dest.ny < dest.n) {
    // this rank's slab: pass B only over the chunks that can reach it
    void *p_list = nullptr;
    if (ws_dev(6, sizeof(int) * (size_t) n0 + 64, &p_list)) return 1;
    int *d_list = (int *) p_list;
    unsigned int *d_count = (unsigned int *) (d_list + n0);
    GX_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), cs));
    need_kernel<<<grid, kThreads, 0, cs>>>(d_streams, ns, n0, d_st[0], dest.n, dest.iy0, dest.ny, d_list, d_count);
    count_launch();
    GX_LAUNCH_CHECK();
    GX_CUDA_CHECK(cudaMemcpyAsync(h_ntail, d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, cs));
    GX_CUDA_CHECK(cudaStreamSynchronize(cs));
    P.list = d_list; P.n_list = (int) *h_ntail;
    grid_b = (P.n_list + kThreads - 1) / kThreads;
  }
  if (grid_b > 0) {
    GX_TIMED("rng_pass_b", cs);
    launch_chunk_kernel<1, T>(grid_b, cs, P, dest);
  }
  count_launch();
  GX_LAUNCH_CHECK();

  // tails: host libm recomputation (the reference's own arithmetic for these draws)
  GX_CUDA_CHECK(cudaMemcpyAsync(h_ntail, d_tail_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, cs));
  GX_CUDA_CHECK(cudaMemcpyAsync(h_end, d_end, sizeof(unsigned long long) * ns, cudaMemcpyDeviceToHost, cs));
  GX_CUDA_CHECK(cudaStreamSynchronize(cs));
  const unsigned int ntail = *h_ntail;
  if (ntail > tail_cap) return fail("gx_noise_draw: tail list overflow");
  for (int s = 0; s < ns; ++s)
    if (reqs[s].n_gauss > 0 && h_end[s] == ~0ull) return fail("gx_noise_draw: raw budget too small for a stream (internal)");
  if (end_raw) end_raw->assign(h_end, h_end + ns);
  if (ntail > 0) {
    void *pin_ev = nullptr, *pin_val = nullptr, *p_val = nullptr;
    if (ws_pin(1, sizeof(TailEvent) * (size_t) ntail, &pin_ev)) return 1;
    if (ws_pin(2, sizeof(double) * (size_t) ntail, &pin_val)) return 1;
    if (ws_dev(5, sizeof(double) * (size_t) ntail, &p_val)) return 1;
    TailEvent *ev = (TailEvent *) pin_ev;
    double *val = (double *) pin_val;
    GX_CUDA_CHECK(cudaMemcpyAsync(ev, d_tails, sizeof(TailEvent) * ntail, cudaMemcpyDeviceToHost, cs));
    GX_CUDA_CHECK(cudaStreamSynchronize(cs));
    host_tail_values(ev, val, ntail);
    GX_CUDA_CHECK(cudaMemcpyAsync(p_val, val, sizeof(double) * ntail, cudaMemcpyHostToDevice, cs));
    GX_TIMED("rng_tail_patch", cs);
    tail_patch_kernel<T><<<(ntail + 127) / 128, 128, 0, cs>>>(d_tails, (const double *) p_val, ntail, d_streams, dest);
    count_launch();
    GX_LAUNCH_CHECK();
    // the pinned staging buffers are reused by the next draw: wait for the upload
    GX_CUDA_CHECK(cudaStreamSynchronize(cs));
  }
  return 0;
}

// ==== staged path for the parallel Fourier order (rng_staged_core.cuh) ====================================================
// Owner table: slot (ix, iy, iz) of the field lives on the rank that holds ky row iy, in a [n][nyl][n/2+1] slab. One GPU:
// world = 1, nyl = n, base[0] = the field. Several GPUs: base[r] are peer-mapped slab pointers (cudaIpc), so the delivery
// kernel of the rank that generated a shell stores every mode straight into its owner's memory over NVLink.
constexpr int kMaxPeers = 8;
template <typename T> struct Owners {
  T *base[kMaxPeers];
  int n, nc, nyl, ny_shift;  // nyl = n / world = 1 << ny_shift
  __device__ __forceinline__ cplx<T> *slot(int a, int b, int c) const {
    const unsigned int ix = (unsigned int) (a < 0 ? a + n : a), iy = (unsigned int) (b < 0 ? b + n : b);
    const unsigned int row = (iy & (unsigned int) (nyl - 1)) + ((ix << ny_shift));   // < 2^24
    return reinterpret_cast<cplx<T> *>(base[iy >> ny_shift]) + ((size_t) row * (unsigned int) nc + (unsigned int) c);
  }
};

struct GenParams {
  const StreamDesc *streams;
  int n_streams;
  int first, count;        // global chunk range of the batch
  const Big *jump;
  ChunkFn *fn;             // per batch chunk
  ListMeta *meta;
  uint32_t *stage;         // kChunkStage words per batch chunk
  StagedTail *tails;
  unsigned int *tail_count;
  unsigned int tail_cap;
  int f64_share;           // of every 8 consecutive CTAs, this many run the generator on the FP64 pipe (rng_core.cuh StateF64)
};

struct SmemColumn {  // column tx of a [slots][kThreads] shared array: conflict-free per-thread storage
  uint32_t (*r)[kThreads];
  int tx;
  __device__ __forceinline__ uint32_t &at(int s) const { return r[s][tx]; }
};

struct DevTailSink {
  StagedTail *tails;
  unsigned int *count;
  unsigned int cap;
  int chunk;
  GX_HD void push(int list, int idx, unsigned int u1, bool positive) {
#ifdef __CUDA_ARCH__
    if (cap == 0) return;
    const unsigned int s = atomicAdd(count, 1u);
    if (s < cap) {
      StagedTail t;
      t.chunk = chunk; t.u1 = u1; t.idx = (unsigned short) idx; t.list = (unsigned char) list; t.positive = positive ? 1 : 0;
      tails[s] = t;
    }
#endif
  }
};

__device__ __forceinline__ int stream_of_chunk(const StreamDesc *streams, int n_streams, int chunk) {
  int lo = 0, hi = n_streams - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (streams[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int MINB, bool FMA>
__global__ void __launch_bounds__(kThreads, MINB) rng_gen_kernel(GenParams P) {
  __shared__ unsigned int s_k[128];
  __shared__ double s_w[128], s_y[128];
  __shared__ uint32_t ring[kRing][kThreads];
  __shared__ uint32_t sbuf[4][kThreads];
  __shared__ uint32_t sinc[kIncCap][kThreads];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { s_k[i] = d_zig_ktab[i]; s_w[i] = d_zig_wtab[i]; s_y[i] = d_zig_ytab[i]; }
  __syncthreads();
  // gen_chunk is warp-collective: the lanes past the end of the batch redo its last chunk (identical stores) and report nothing
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = slot < P.count;
  const int rel = in_range ? slot : P.count - 1;
  const int chunk = P.first + rel;
  const StreamDesc &sd = P.streams[stream_of_chunk(P.streams, P.n_streams, chunk)];
  const int q = chunk - sd.first_chunk;
  const State st = chunk_start_state(sd, q, P.jump);
  DevTailSink sink{P.tails, P.tail_count, in_range ? P.tail_cap : 0u, rel};
  ChunkFn f;
  ListMeta m;
  if ((int) (blockIdx.x & 7u) < P.f64_share)
    gen_chunk<2>(sd, q == 0, st, SmemColumn{ring, (int) threadIdx.x}, SmemColumn{sbuf, (int) threadIdx.x},
                 SmemColumn{sinc, (int) threadIdx.x}, s_k, s_w, s_y, P.stage + (size_t) rel * kChunkStage, f, m, sink);
  else
    gen_chunk<FMA ? 1 : 0>(sd, q == 0, st, SmemColumn{ring, (int) threadIdx.x}, SmemColumn{sbuf, (int) threadIdx.x},
                           SmemColumn{sinc, (int) threadIdx.x}, s_k, s_w, s_y, P.stage + (size_t) rel * kChunkStage, f, m, sink);
  if (in_range) { P.fn[rel] = f; P.meta[rel] = m; }
}

__global__ void resolve_kernel(const ChunkFn *fn, const ScanState *scan, const ListMeta *meta, uint32_t *stage, ChunkLoc *loc,
                               int count) {
  const int rel = blockIdx.x * blockDim.x + threadIdx.x;
  if (rel >= count) return;
  const bool first = fn[rel].reset != 0;
  const ScanState ss = scan[rel];
  loc[rel] = resolve_chunk(meta[rel], first ? 0 : ss.entry, first ? 0 : ss.base, stage + (size_t) rel * kChunkStage);
}

// staged tails of a batch -> (stream, gaussian index) events of the draws that lie on the chunk's true path
__global__ void tail_resolve_kernel(const StagedTail *st, const unsigned int *count, unsigned int cap, const ChunkFn *fn,
                                    const ScanState *scan, const ListMeta *meta, const ChunkLoc *loc, const StreamDesc *streams,
                                    int n_streams, int first_chunk, TailEvent *out, unsigned int *out_count, unsigned int out_cap,
                                    unsigned int *overflow) {
  if (*count > cap && blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1u;
  const unsigned int n = min(*count, cap);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const StagedTail t = st[i];
    const bool first = fn[t.chunk].reset != 0;
    const int pi = path_index(meta[t.chunk], first ? 0 : scan[t.chunk].entry, t.list, t.idx);
    if (pi < 0) continue;
    const int sid = stream_of_chunk(streams, n_streams, first_chunk + t.chunk);
    const long long g = loc[t.chunk].base + pi;
    if (g >= streams[sid].n_gauss) continue;
    const unsigned int s = atomicAdd(out_count, 1u);
    if (s < out_cap) { TailEvent ev; ev.g = g; ev.stream = sid; ev.u1 = t.u1; ev.positive = t.positive; out[s] = ev; }
  }
}

__global__ void shell_scale_kernel(const ChunkLoc *loc, const int *shell_first, int chunk0, int sa, int sb, float *scale) {
  const int ks = sa + blockIdx.x * blockDim.x + threadIdx.x;
  if (ks >= sb) return;
  scale[ks] = locate_scale(loc, shell_first[ks] - chunk0, shell_first[ks + 1] - shell_first[ks]);
}

template <typename T> struct DeliverParams {
  Owners<T> own;
  int reverse;
  double sigma;
  const T *wtab;
  const int *shell_first;  // [n/2 + 1]: global index of the first chunk of shell ks (entry ks + 1: its end)
  const float *shell_scale; // [n/2]: locate_scale of shell ks (valid for the batch's shells)
  int chunk0;              // global index of the batch's first chunk
  const ChunkLoc *loc;     // per batch chunk
  const uint32_t *stage;
  int sa, sb;              // shells [sa, sb) of the batch
  int p_origin, q_origin;  // (p, q) of item 0, lane 0
  int n_qblocks, n_items;  // items = (p, block of 32 q values), q-block fastest
  int item_shift;          // rotation of the item order (rank * n_items / world)
};

// CTA = one k1 value p and 32 consecutive k2 values (lanes); the 8 warps take 8 consecutive shells per iteration.
template <typename T>
__global__ void __launch_bounds__(256) deliver_kernel(DeliverParams<T> P) {
  __shared__ cplx<T> tile[32][9];
  __shared__ unsigned int vmask[8];
  __shared__ double s_w[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 128) s_w[tid] = d_zig_wtab[tid];
  __syncthreads();
  // work items (p, q-block): one CTA each, or -- when the grid is smaller than the item count (several GPUs: the kernel is
  // NVLink-bound and leaves most of every SM to the generation of the next batch) -- a grid-stride walk over them
  for (int it0 = blockIdx.x; it0 < P.n_items; it0 += gridDim.x) {
  // (several ranks: each starts its sweep over k1 at a different place, so that at any moment the ranks store into
  // different owners -- the face with ky = k1 sends a whole CTA's output to one rank)
  const int item = it0 + P.item_shift < P.n_items ? it0 + P.item_shift : it0 + P.item_shift - P.n_items;
  const int p = P.p_origin + item / P.n_qblocks;
  const int q0 = P.q_origin + 32 * (item % P.n_qblocks);
  const int q = q0 + lane;
  const int need_p = max(-p, p + 1), need_q = max(-q, q + 1);      // first shell whose loops reach p / q
  const int need_blk = max(need_p, q0 >= 0 ? q0 + 1 : -(q0 + 31));  // ... reach any lane of this block
  const int start = max(P.sa, need_blk) & ~7;
  for (int ks0 = start; ks0 < P.sb; ks0 += 8) {
    const int ks = ks0 + warp;
    const bool valid = ks >= P.sa && ks < P.sb && ks >= need_p && ks >= need_q;
    bool w2 = false;
    if (valid) {
      const int c0 = ((p + ks) * 2 * ks + (q + ks)) * 3;
      const long long g0 = 2ll * c0;
      const int fc = P.shell_first[ks] - P.chunk0, nch = P.shell_first[ks + 1] - P.shell_first[ks];
      int ci = locate_chunk(P.loc, fc, nch, g0, P.shell_scale[ks]);
      ChunkLoc L = P.loc[ci];
      long long hi = (ci + 1 < fc + nch) ? P.loc[ci + 1].base : 0x7fffffffffffffffll;
      double d[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const long long g = g0 + k;
        if (g >= hi) {  // the triple straddles a chunk boundary
          ++ci;
          L = P.loc[ci];
          hi = (ci + 1 < fc + nch) ? P.loc[ci + 1].base : 0x7fffffffffffffffll;
        }
        d[k] = unpack_draw(fetch_draw(L, P.stage + (size_t) ci * kChunkStage, (int) (g - L.base)), s_w);
      }
#pragma unroll
      for (int face = 0; face < 3; ++face) {
        const FaceSlot fs = face_slot(ks, p, q, face);
        if (!fs.write) continue;
        T re, im;
        call_value<T>(d[2 * face], d[2 * face + 1], P.sigma, P.wtab, fs.m, P.reverse, re, im);
        if (face < 2) {  // kz = +-q: the warp's stores are one contiguous run of a row
          *P.own.slot(fs.a, fs.b, fs.c) = mk<T>(re, fs.neg ? -im : im);
          if (fs.second) *P.own.slot(-fs.a, -fs.b, 0) = mk<T>(re, -im);
        } else {         // kz = ks: 32 different rows -- through the tile, 8 shells side by side
          tile[lane][warp] = mk<T>(re, im);
          w2 = true;
        }
      }
    }
    const unsigned int wm = __ballot_sync(0xffffffffu, w2);
    if (lane == 0) vmask[warp] = wm;
    __syncthreads();
    {
      const int qq = tid >> 3, kk = tid & 7;
      if ((vmask[kk] >> qq) & 1u) *P.own.slot(p, q0 + qq, ks0 + kk) = tile[qq][kk];
    }
    __syncthreads();
  }
  }
}

template <typename T>
__global__ void staged_patch_kernel(const TailEvent *ev, const double *values, unsigned int n, const StreamDesc *streams,
                                    Owners<T> own, int reverse, double sigma, const T *wtab) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long g = ev[i].g;
  const int ks = streams[ev[i].stream].shell;
  const long long c = g >> 1;
  const int comp = (int) (g & 1), face = (int) (c % 3);
  const long long t = c / 3;
  const int p = (int) (t / (2 * ks)) - ks, q = (int) (t % (2 * ks)) - ks;
  const FaceSlot fs = face_slot(ks, p, q, face);
  if (!fs.write) return;
  T v = (T) GX_DMUL(sigma, values[i]);
  if (wtab) v = (T) GX_DMUL(v, __ldg(wtab + fs.m));
  const int part = reverse ? 1 - comp : comp;
  const T vconj = part == 1 ? -v : v;
  reinterpret_cast<T *>(own.slot(fs.a, fs.b, fs.c))[part] = fs.neg ? vconj : v;
  if (fs.second) reinterpret_cast<T *>(own.slot(-fs.a, -fs.b, 0))[part] = vconj;
}

// the slots no call writes, in this rank's rows [iy0, iy0 + nyl): one thread per (kx, ky) row
template <typename T> __global__ void zero_undrawn_kernel(cplx<T> *slab, int n, int iy0, int nyl) {
  const long row = blockIdx.x * (long) blockDim.x + threadIdx.x;
  if (row >= (long) n * nyl) return;
  const int h = n >> 1, nc = h + 1;
  const int ix = (int) (row / nyl), iy = iy0 + (int) (row % nyl);
  cplx<T> *r = slab + (size_t) row * nc;
  const cplx<T> z = mk<T>((T) 0, (T) 0);
  if (ix == h || iy == h) { for (int c = 0; c < nc; ++c) r[c] = z; return; }
  r[h] = z;
  const int a = ix <= h ? ix : ix - n, b = iy <= h ? iy : iy - n;
  const int A = max(abs(a), abs(b));
  if (never_drawn(n, ix, iy, A)) r[A] = z;
  if (A > 0 && never_drawn(n, ix, iy, 0)) for (int c = 0; c < A; ++c) r[c] = z;  // kz < A: the same status along the row
}

struct Shard { int world = 1, rank = 0; void *peers[kMaxPeers] = {nullptr}; };

static int staged_minb() {
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("GX_RNG_GEN_MINB"); minb = e ? atoi(e) : 5; }
  return minb;
}

// Parallel Fourier order through the staged path. `field`: the whole field (world 1) or this rank's ky-slab; `scratch`:
// device memory for the staged draws (any size from one shell's worth up; nullptr: allocated here for the call).
template <typename T>
static int noise_parallel_staged(T *field, int n, unsigned long seed, int reverse, const T *wtab, void *scratch,
                                 size_t scratch_bytes, cudaStream_t cs, const Shard &shard) {
  if (ensure_jump_table()) return 1;
  if (ws_bind_device()) return 1;
  const int h = n / 2, ns = h - 1;
  const int world = shard.world, rank = shard.rank, nyl = n / world;
  Owners<T> own;
  own.n = n; own.nc = h + 1; own.nyl = nyl; own.ny_shift = 0;
  while ((1 << own.ny_shift) < nyl) ++own.ny_shift;
  if ((1 << own.ny_shift) != nyl) return fail("gx_noise_draw: n / world must be a power of two");
  for (int r = 0; r < kMaxPeers; ++r) own.base[r] = nullptr;
  if (world == 1) own.base[0] = field;
  else for (int r = 0; r < world; ++r) own.base[r] = (T *) (r == rank ? (void *) field : shard.peers[r]);
  {
    const long rows = (long) n * nyl;
    zero_undrawn_kernel<T><<<(unsigned) ((rows + 255) / 256), 256, 0, cs>>>(reinterpret_cast<cplx<T> *>(field), n, rank * nyl, nyl);
    count_launch();
    GX_LAUNCH_CHECK();
  }
  if (ns <= 0) return 0;
  // streams: one per shell, seed + ks (randomfieldgenerator.hpp:228-229), 24 ks^2 gaussians each
  void *pin0 = nullptr;
  if (ws_pin(0, sizeof(StreamDesc) * ns + sizeof(int) * (h + 2) + 64, &pin0)) return 1;
  StreamDesc *sd = (StreamDesc *) pin0;
  int *shell_first = (int *) ((char *) pin0 + sizeof(StreamDesc) * ns);  // [ks], ks = 0..h
  unsigned int *h_ntail = (unsigned int *) (shell_first + h + 1);  // 3 words
  long long n_chunks = 0, total_gauss = 0;
  shell_first[0] = 0;
  for (int s = 0; s < ns; ++s) {
    StreamDesc &d = sd[s];
    const int ks = s + 1;
    d.s0 = seed_state(seed + (unsigned long) ks);
    d.b0 = 0; d.entry0 = 0; d.n_gauss = 24ll * ks * ks; d.shell = ks;
    d.xs = state_to_big(d.s0);
    d.first_chunk = (int) n_chunks;
    shell_first[ks] = (int) n_chunks;
    n_chunks += (raw_budget(d.n_gauss) + kChunkLen - 1) / kChunkLen;
    total_gauss += d.n_gauss;
  }
  shell_first[h] = (int) n_chunks;
  // this rank's shells: contiguous, balanced by chunk count (generator work)
  int s_lo = 1, s_hi = h;
  if (world > 1) {
    auto cut = [&](int r) {
      const long long target = n_chunks * r / world;
      int ks = 1;
      while (ks < h && shell_first[ks] < target) ++ks;
      return ks;
    };
    s_lo = cut(rank);
    s_hi = rank + 1 == world ? h : cut(rank + 1);
  }
  if (s_lo >= s_hi) return 0;
  const long long my_chunks = shell_first[s_hi] - shell_first[s_lo];
  const long long largest = shell_first[s_hi] - shell_first[s_hi - 1];
  const size_t per_chunk = (size_t) kChunkStage * sizeof(uint32_t);
  void *own_scratch = nullptr;
  if (!scratch || scratch_bytes < (size_t) largest * per_chunk) {
    size_t free_b = 0, total_b = 0;
    GX_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    size_t want = (size_t) my_chunks * per_chunk;
    size_t cap = free_b / 2;
    if (const char *e = getenv("GX_NOISE_STAGE_BYTES")) cap = (size_t) atoll(e);  // test hook: forces several batches
    cap = std::max((size_t) largest * per_chunk, cap);
    if (want > cap) want = cap;
    GX_CUDA_CHECK(cudaMalloc(&own_scratch, want));
    scratch = own_scratch; scratch_bytes = want;
  }
  // Several ranks: the delivery stores over NVLink and is bound by it, the generation is pure integer work. With two halves of
  // the scratch buffer and of the per-batch arrays, batch b + 1 is generated (stream s_gen) while batch b is delivered (s_del).
  static int pipe_env = -1;
  // (measured on 2 B200s at 2048^3, profiles/r2s10_*: no gain -- the delivery needs most of the GPU's warps to keep NVLink busy,
  // with 1-2 CTAs per SM it drops from 32 to 48-76 ms -- so this is off unless GX_NOISE_PIPELINE=1)
  if (pipe_env < 0) { const char *e = getenv("GX_NOISE_PIPELINE"); pipe_env = (e && e[0] == '1') ? 1 : 0; }
  const bool pipelined = world > 1 && pipe_env && scratch_bytes / 2 >= (size_t) largest * per_chunk;
  const int nslots = pipelined ? 2 : 1;
  const size_t slot_bytes = (scratch_bytes / nslots) & ~(size_t) 255;
  long long cap_chunks = (long long) (slot_bytes / per_chunk);
  {
    // equal batches: with the capacity as the only limit a 2048^3 draw was two full batches and a 2 % one, whose kernels ran
    // on a nearly empty GPU
    const long long nb = (my_chunks + cap_chunks - 1) / cap_chunks;
    const long long even = (my_chunks + nb - 1) / nb + largest;   // whole shells: allow one shell of slack
    if (even < cap_chunks) cap_chunks = even;
  }
  // batches of whole shells
  std::vector<std::pair<int, int>> batches;
  long long max_batch = 0, max_batch_gauss = 0;
  for (int a = s_lo; a < s_hi;) {
    int b = a + 1;
    while (b < s_hi && shell_first[b + 1] - shell_first[a] <= cap_chunks) ++b;
    batches.push_back({a, b});
    max_batch = std::max<long long>(max_batch, shell_first[b] - shell_first[a]);
    long long gsum = 0;
    for (int ks = a; ks < b; ++ks) gsum += 24ll * ks * ks;
    max_batch_gauss = std::max(max_batch_gauss, gsum);
    a = b;
  }
  // per-batch arrays
  std::vector<int> level_n;
  level_n.push_back((int) max_batch);
  while (level_n.back() > 4096) level_n.push_back((level_n.back() + 63) / 64);
  size_t fn_total = 0;
  for (int v : level_n) fn_total += (size_t) v;
  long long my_gauss = 0;
  for (int ks = s_lo; ks < s_hi; ++ks) my_gauss += 24ll * ks * ks;
  const unsigned int staged_cap = (unsigned int) std::min<long long>((long long) (max_batch_gauss * 1.5e-3) + 8192, 600000000ll);
  const unsigned int tail_cap = (unsigned int) std::min<long long>((long long) (my_gauss * 8e-4) + 4096, 400000000ll);
  void *p_streams, *p_misc, *p_tails, *p_fn, *p_st, *p_meta, *p_staged;
  if (ws_dev(0, sizeof(StreamDesc) * ns + sizeof(int) * (h + 1) + sizeof(float) * (h + 1), &p_streams)) return 1;
  const size_t meta_bytes = ((sizeof(ListMeta) + sizeof(ChunkLoc)) * (size_t) max_batch + 255) & ~(size_t) 255;
  if (ws_dev(1, 64, &p_misc)) return 1;
  if (ws_dev(2, sizeof(TailEvent) * (size_t) tail_cap, &p_tails)) return 1;
  if (ws_dev(3, sizeof(ChunkFn) * fn_total * nslots, &p_fn)) return 1;
  if (ws_dev(4, sizeof(ScanState) * fn_total * nslots, &p_st)) return 1;
  if (ws_dev(6, meta_bytes * nslots, &p_meta)) return 1;
  if (ws_dev(7, sizeof(StagedTail) * (size_t) staged_cap * nslots, &p_staged)) return 1;
  StreamDesc *d_streams = (StreamDesc *) p_streams;
  int *d_shell_first = (int *) ((char *) p_streams + sizeof(StreamDesc) * ns);
  float *d_shell_scale = (float *) (d_shell_first + h + 1);
  // p_misc: [0] final tail count, [1] staged-overflow flag, [2], [3] staged tail counts of the two slots
  unsigned int *d_tail_count = (unsigned int *) p_misc, *d_staged_counts = d_tail_count + 2;
  GX_CUDA_CHECK(cudaMemcpyAsync(d_streams, sd, sizeof(StreamDesc) * ns + sizeof(int) * (h + 1), cudaMemcpyHostToDevice, cs));
  GX_CUDA_CHECK(cudaMemsetAsync(d_tail_count, 0, 4 * sizeof(unsigned int), cs));
  const int nl = (int) level_n.size();
  std::vector<ChunkFn *> d_fn(nl);
  std::vector<ScanState *> d_st(nl);
  const double sigma = (double) (T) (1.0 / sqrt(2.0));

  static cudaEvent_t ev_tails = nullptr, ev_vals = nullptr;
  static cudaStream_t s_tail = nullptr;
  static cudaStream_t s_gen = nullptr, s_del = nullptr;
  static cudaEvent_t ev_ready[2], ev_free[2], ev_fork, ev_join[2];
  static int s_dev = -1;
  if (pipelined) {
    int dev = 0;
    GX_CUDA_CHECK(cudaGetDevice(&dev));
    if (s_dev != dev) {
      GX_CUDA_CHECK(cudaStreamCreateWithFlags(&s_gen, cudaStreamNonBlocking));
      GX_CUDA_CHECK(cudaStreamCreateWithFlags(&s_del, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming));
        GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
        GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
      }
      GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      s_dev = dev;
    }
    GX_CUDA_CHECK(cudaEventRecord(ev_fork, cs));
    GX_CUDA_CHECK(cudaStreamWaitEvent(s_gen, ev_fork, 0));
    GX_CUDA_CHECK(cudaStreamWaitEvent(s_del, ev_fork, 0));
  }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }

  for (size_t bi = 0; bi < batches.size(); ++bi) {
    const auto &bt = batches[bi];
    const int slot = pipelined ? (int) (bi & 1) : 0;
    const cudaStream_t cg = pipelined ? s_gen : cs, cd = pipelined ? s_del : cs;
    void *const scratch_b = (char *) scratch + (size_t) slot * slot_bytes;
    ChunkFn *const fn_b = (ChunkFn *) p_fn + (size_t) slot * fn_total;
    ScanState *const st_b = (ScanState *) p_st + (size_t) slot * fn_total;
    ChunkLoc *const d_loc = (ChunkLoc *) ((char *) p_meta + (size_t) slot * meta_bytes);
    ListMeta *const d_meta = (ListMeta *) ((char *) d_loc + sizeof(ChunkLoc) * (size_t) max_batch);
    StagedTail *const staged_b = (StagedTail *) p_staged + (size_t) slot * staged_cap;
    unsigned int *const d_staged_count = d_staged_counts + slot;
    if (pipelined && bi >= 2) GX_CUDA_CHECK(cudaStreamWaitEvent(cg, ev_free[slot], 0));  // the slot's previous batch is delivered
    const int sa = bt.first, sb = bt.second;
    const int c0 = shell_first[sa], cnt = shell_first[sb] - c0;
    std::vector<int> ln;
    ln.push_back(cnt);
    while (ln.back() > 4096) ln.push_back((ln.back() + 63) / 64);
    const int bl = (int) ln.size();
    {
      size_t off = 0;
      for (int l = 0; l < bl; ++l) { d_fn[l] = fn_b + off; d_st[l] = st_b + off; off += (size_t) ln[l]; }
    }
    GX_CUDA_CHECK(cudaMemsetAsync(d_staged_count, 0, sizeof(unsigned int), cg));
    GenParams G;
    G.streams = d_streams; G.n_streams = ns; G.first = c0; G.count = cnt; G.jump = g_jump; G.fn = d_fn[0]; G.meta = d_meta;
    G.stage = (uint32_t *) scratch_b; G.tails = staged_b; G.tail_count = d_staged_count; G.tail_cap = staged_cap;
    {
      static int share = -1;
      if (share < 0) { const char *e = getenv("GX_RNG_F64_SHARE"); share = e ? atoi(e) : 0; }
      G.f64_share = share;
    }
    {
      GX_TIMED("rng_gen", cg);
      const int grid = (cnt + kThreads - 1) / kThreads;
      const int mb = staged_minb();
      // mask of the SWB step on the fma pipe (IMAD.HI + IMAD) instead of the alu pipe: measured SLOWER at 2048^3
      // (100 against 87 ms, profiles/r2s4_*): off unless GX_RNG_FMA_MASK=1
      static int fma = -1;
      if (fma < 0) { const char *e = getenv("GX_RNG_FMA_MASK"); fma = (e && e[0] == '1') ? 1 : 0; }
      if (fma) {
        if (mb == 5) rng_gen_kernel<5, true><<<grid, kThreads, 0, cg>>>(G);
        else if (mb == 8) rng_gen_kernel<8, true><<<grid, kThreads, 0, cg>>>(G);
        else rng_gen_kernel<6, true><<<grid, kThreads, 0, cg>>>(G);
      } else {
        if (mb == 5) rng_gen_kernel<5, false><<<grid, kThreads, 0, cg>>>(G);
        else if (mb == 8) rng_gen_kernel<8, false><<<grid, kThreads, 0, cg>>>(G);
        else rng_gen_kernel<6, false><<<grid, kThreads, 0, cg>>>(G);
      }
      count_launch();
    }
    GX_LAUNCH_CHECK();
    {
      GX_TIMED("rng_scan", cg);
      for (int l = 1; l < bl; ++l) {
        scan_up_kernel<<<(ln[l] + 127) / 128, 128, 0, cg>>>(d_fn[l - 1], ln[l - 1], d_fn[l], ln[l]);
        count_launch();
      }
      scan_top_kernel<<<1, 32, 0, cg>>>(d_fn[bl - 1], ln[bl - 1], d_st[bl - 1]);
      count_launch();
      for (int l = bl - 1; l >= 1; --l) {
        scan_down_kernel<<<(ln[l] + 127) / 128, 128, 0, cg>>>(d_fn[l - 1], ln[l - 1], d_st[l], ln[l], d_st[l - 1]);
        count_launch();
      }
      resolve_kernel<<<(cnt + 127) / 128, 128, 0, cg>>>(d_fn[0], d_st[0], d_meta, (uint32_t *) scratch_b, d_loc, cnt);
      count_launch();
      shell_scale_kernel<<<(sb - sa + 127) / 128, 128, 0, cg>>>(d_loc, d_shell_first, c0, sa, sb, d_shell_scale);
      count_launch();
      tail_resolve_kernel<<<128, 256, 0, cg>>>((const StagedTail *) staged_b, d_staged_count, staged_cap, d_fn[0], d_st[0], d_meta,
                                               d_loc, d_streams, ns, c0, (TailEvent *) p_tails, d_tail_count, tail_cap,
                                               d_tail_count + 1);
      count_launch();
    }
    GX_LAUNCH_CHECK();
    if (bi + 1 == batches.size()) {  // every tail event of the draw is listed: the host side can start under the last delivery
      if (!ev_tails) GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_tails, cudaEventDisableTiming));
      GX_CUDA_CHECK(cudaEventRecord(ev_tails, cg));
    }
    DeliverParams<T> D;
    D.own = own; D.reverse = reverse; D.sigma = sigma; D.wtab = wtab; D.shell_first = d_shell_first; D.shell_scale = d_shell_scale; D.chunk0 = c0;
    D.loc = d_loc; D.stage = (const uint32_t *) scratch_b; D.sa = sa; D.sb = sb;
    const int top = sb - 1;                       // p, q in [-top, top)
    D.p_origin = -top;
    D.q_origin = -((top + 31) / 32) * 32;
    D.n_qblocks = (top - D.q_origin + 31) / 32;
    D.n_items = D.n_qblocks * 2 * top;
    D.item_shift = world > 1 ? (int) ((long long) D.n_items * rank / world) : 0;
    if (pipelined) {
      GX_CUDA_CHECK(cudaEventRecord(ev_ready[slot], cg));
      GX_CUDA_CHECK(cudaStreamWaitEvent(cd, ev_ready[slot], 0));
    }
    {
      static int per_sm = -1;  // CTAs per SM of the NVLink-bound delivery when it shares the GPU with the generation
      if (per_sm < 0) { const char *e = getenv("GX_DELIVER_CTAS_PER_SM"); per_sm = e ? atoi(e) : 1; }
      const int grid = pipelined ? std::min(D.n_items, std::max(1, per_sm) * sms) : D.n_items;
      GX_TIMED("rng_deliver", cd);
      deliver_kernel<T><<<grid, 256, 0, cd>>>(D);
      count_launch();
    }
    GX_LAUNCH_CHECK();
    if (pipelined) GX_CUDA_CHECK(cudaEventRecord(ev_free[slot], cd));
  }
  if (pipelined) {
    GX_CUDA_CHECK(cudaEventRecord(ev_join[0], s_gen));
    GX_CUDA_CHECK(cudaEventRecord(ev_join[1], s_del));
    GX_CUDA_CHECK(cudaStreamWaitEvent(cs, ev_join[0], 0));
    GX_CUDA_CHECK(cudaStreamWaitEvent(cs, ev_join[1], 0));
  }

  // tails: host libm recomputation (the reference's own arithmetic for these draws), then patched over the delivered values.
  // The event list is complete once the last batch is resolved, so its copy to the host, the log() calls and the upload of
  // the values run on a side stream UNDER the last delivery kernel (17 ms at 2048^3) instead of on an idle GPU after it.
  if (!s_tail) {
    GX_CUDA_CHECK(cudaStreamCreateWithFlags(&s_tail, cudaStreamNonBlocking));
    GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_vals, cudaEventDisableTiming));
  }
  GX_CUDA_CHECK(cudaStreamWaitEvent(s_tail, ev_tails, 0));
  GX_CUDA_CHECK(cudaMemcpyAsync(h_ntail, d_tail_count, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s_tail));
  GX_CUDA_CHECK(cudaStreamSynchronize(s_tail));
  const unsigned int ntail = h_ntail[0];
  int rc = 0;
  if (h_ntail[1]) rc = fail("gx_noise_draw: staged tail list overflow");
  else if (ntail > tail_cap) rc = fail("gx_noise_draw: tail list overflow");
  if (!rc && ntail > 0) {
    void *pin_ev = nullptr, *pin_val = nullptr, *p_val = nullptr;
    if (ws_pin(1, sizeof(TailEvent) * (size_t) ntail, &pin_ev)) return 1;
    if (ws_pin(2, sizeof(double) * (size_t) ntail, &pin_val)) return 1;
    if (ws_dev(5, sizeof(double) * (size_t) ntail, &p_val)) return 1;
    TailEvent *ev = (TailEvent *) pin_ev;
    double *val = (double *) pin_val;
    GX_CUDA_CHECK(cudaMemcpyAsync(ev, p_tails, sizeof(TailEvent) * ntail, cudaMemcpyDeviceToHost, s_tail));
    GX_CUDA_CHECK(cudaStreamSynchronize(s_tail));
    host_tail_values(ev, val, ntail);
    GX_CUDA_CHECK(cudaMemcpyAsync(p_val, val, sizeof(double) * ntail, cudaMemcpyHostToDevice, s_tail));
    GX_CUDA_CHECK(cudaEventRecord(ev_vals, s_tail));
    GX_CUDA_CHECK(cudaStreamWaitEvent(cs, ev_vals, 0));
    GX_TIMED("rng_tail_patch", cs);
    staged_patch_kernel<T><<<(ntail + 127) / 128, 128, 0, cs>>>((const TailEvent *) p_tails, (const double *) p_val, ntail, d_streams,
                                                             own, reverse, sigma, wtab);
    count_launch();
    GX_LAUNCH_CHECK();
  }
  GX_CUDA_CHECK(cudaStreamSynchronize(cs));  // the pinned staging buffers and the scratch are reused by the next draw
  if (own_scratch) GX_CUDA_CHECK(cudaFree(own_scratch));
  return rc;
}

// world == 1: the whole field. world > 1: rank's slab only -- Fourier orders fill the ky-slab [n][n/world][nc] (every
// rank walks every stream and keeps the draws it owns), the real order fills the x-slab and stays in real space (the
// caller runs the slab-decomposed r2c and zeroes Re(DC)).
static bool staged_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("GX_NOISE_STAGED"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

template <typename T>
static int noise_draw_t(void *field, int n, unsigned long seed, int mode, int reverse, unsigned long long *offset_io,
                        cudaStream_t cs, int world = 1, int rank = 0, Phase phase = Phase(), const void *wtab = nullptr,
                        void *scratch = nullptr, size_t scratch_bytes = 0) {
  if (mode == GX_NOISE_FOURIER_PARALLEL && world == 1 && phase.which == 0 && staged_enabled())
    return noise_parallel_staged<T>((T *) field, n, seed, reverse, (const T *) wtab, scratch, scratch_bytes, cs, Shard());
  const size_t bytes = (size_t) 2 * n * n * (n / 2 + 1) * sizeof(T) / (size_t) world;
  if (phase.which != 1) GX_CUDA_CHECK(cudaMemsetAsync(field, 0, bytes, cs));
  Dest<T> dest;
  dest.n = n; dest.reverse = reverse; dest.field = (T *) field; dest.list = nullptr;
  dest.ny = n / world; dest.iy0 = rank * dest.ny;
  dest.nrows = (long long) n * n / world; dest.row0 = rank * dest.nrows;
  dest.sigma = (double) (T) (1.0 / sqrt(2.0));
  dest.wtab = (const T *) wtab;
  if (wtab && mode == GX_NOISE_REAL) return fail("noise_draw_scaled: the real-space order has no per-mode factor");
  const long long n3 = (long long) n * n * n;
  const unsigned long long start = offset_io ? *offset_io : 0ull;
  std::vector<StreamReq> reqs;
  std::vector<unsigned long long> ends;
  if (mode == GX_NOISE_REAL) {
    dest.mode = MODE_REAL;
    reqs.push_back({seed, start, n3, 0});
  } else if (mode == GX_NOISE_FOURIER_SERIAL) {
    dest.mode = MODE_FOURIER_SERIAL;
    const long long h = n / 2;
    reqs.push_back({seed, start, 2 * calls_before_shell_host(h), 0});
  } else if (mode == GX_NOISE_FOURIER_PARALLEL) {
    dest.mode = MODE_FOURIER_PARALLEL;
    for (int ks = 1; ks < n / 2; ++ks) reqs.push_back({seed + (unsigned long) ks, 0ull, 24ll * ks * ks, ks});
  } else {
    return fail("gx_noise_draw: unknown mode");
  }
  int rc = run_streams<T>(reqs, dest, &ends, cs, phase);
  if (rc) return rc;
  if (phase.which == 1) return 0;
  if (mode == GX_NOISE_REAL && world == 1) {
    rc = fft::fft3d_any<T>((T *) field, n, -1, cs);
    if (rc) return rc;
    zero_dc_real_kernel<T><<<1, 1, 0, cs>>>((T *) field);  // tools::set_zero(fieldData[0]) :195
    count_launch();
    GX_LAUNCH_CHECK();
  }
  if (offset_io && mode != GX_NOISE_FOURIER_PARALLEL && !ends.empty()) *offset_io = ends[0];
  return 0;
}

}  // namespace rng

// the real-space order (N^3 sequential draws, then the r2c) takes any even size; the Fourier-space orders are built on
// power-of-two shells
static int check_noise_n(int n, int mode) {
  if (n < 4 || n > 4096 || (n & 1)) return fail("grid size must be even and in [4,4096]");
  if (!is_pow2(n) && mode != GX_NOISE_REAL) return fail("the Fourier-space draw orders need a power-of-two grid size");
  return 0;
}

int noise_draw_scaled(void *field, int n, int dtype, unsigned long seed, int mode, int reverse_order,
                      unsigned long long *stream_offset_io, cudaStream_t st, const void *table, void *scratch,
                      size_t scratch_bytes) {
  if (check_noise_n(n, mode)) return 1;
  if (dtype == GX_F64) return rng::noise_draw_t<double>(field, n, seed, mode, reverse_order, stream_offset_io, st, 1, 0, rng::Phase(), table, scratch, scratch_bytes);
  if (dtype == GX_F32) return rng::noise_draw_t<float>(field, n, seed, mode, reverse_order, stream_offset_io, st, 1, 0, rng::Phase(), table, scratch, scratch_bytes);
  return fail("bad dtype");
}
}  // namespace gx

using namespace gx;
using namespace gx::rng;

extern "C" {

int gx_noise_draw(void *field, int n, int dtype, unsigned long seed, int mode, int reverse_order,
                  unsigned long long *stream_offset_io, void *stream) {
  if (gx::check_noise_n(n, mode)) return 1;
  if (dtype == GX_F64) return noise_draw_t<double>(field, n, seed, mode, reverse_order, stream_offset_io, (cudaStream_t) stream);
  if (dtype == GX_F32) return noise_draw_t<float>(field, n, seed, mode, reverse_order, stream_offset_io, (cudaStream_t) stream);
  return gx::fail("bad dtype");
}

int gx_slab_noise_draw(void *slab, int n, int dtype, int world, int rank, unsigned long seed, int mode,
                       int reverse_order, unsigned long long *stream_offset_io, void *stream) {
  if (!gx::is_pow2(n) || n < 4 || n > 4096) return gx::fail("grid size must be a power of two in [4,4096]");
  if (world < 1 || n % world != 0 || rank < 0 || rank >= world) return gx::fail("gx_slab_noise_draw: bad world/rank");
  if (dtype == GX_F64) return noise_draw_t<double>(slab, n, seed, mode, reverse_order, stream_offset_io, (cudaStream_t) stream, world, rank);
  if (dtype == GX_F32) return noise_draw_t<float>(slab, n, seed, mode, reverse_order, stream_offset_io, (cudaStream_t) stream, world, rank);
  return gx::fail("bad dtype");
}

// chunk functions of the per-shell streams: sum over shells of ceil(raw_budget(24 ks^2) / kChunkLen), split between ranks
size_t gx_slab_noise_fn_bytes(int n, int world) {
  long long chunks = 0;
  for (int ks = 1; ks < n / 2; ++ks) chunks += (raw_budget(24ll * ks * ks) + kChunkLen - 1) / kChunkLen;
  if (world < 1) world = 1;
  return (size_t) ((chunks + world - 1) / world) * sizeof(ChunkFn);
}

int gx_slab_noise_pass_a(void *fn_local, int n, int dtype, int world, int rank, unsigned long seed, void *stream) {
  if (!gx::is_pow2(n) || n < 4 || n > 4096) return gx::fail("grid size must be a power of two in [4,4096]");
  if (world < 1 || n % world != 0 || rank < 0 || rank >= world) return gx::fail("gx_slab_noise_pass_a: bad world/rank");
  Phase ph; ph.which = 1; ph.world = world; ph.rank = rank; ph.fn_io = fn_local;
  if (dtype == GX_F64) return noise_draw_t<double>(nullptr, n, seed, GX_NOISE_FOURIER_PARALLEL, 0, nullptr, (cudaStream_t) stream, world, rank, ph);
  if (dtype == GX_F32) return noise_draw_t<float>(nullptr, n, seed, GX_NOISE_FOURIER_PARALLEL, 0, nullptr, (cudaStream_t) stream, world, rank, ph);
  return gx::fail("bad dtype");
}

int gx_slab_noise_finish(void *slab, const void *fn_all, int n, int dtype, int world, int rank, unsigned long seed,
                         int reverse_order, void *stream) {
  if (!gx::is_pow2(n) || n < 4 || n > 4096) return gx::fail("grid size must be a power of two in [4,4096]");
  if (world < 1 || n % world != 0 || rank < 0 || rank >= world) return gx::fail("gx_slab_noise_finish: bad world/rank");
  Phase ph; ph.which = 2; ph.world = world; ph.rank = rank; ph.fn_io = const_cast<void *>(fn_all);
  if (dtype == GX_F64) return noise_draw_t<double>(slab, n, seed, GX_NOISE_FOURIER_PARALLEL, reverse_order, nullptr, (cudaStream_t) stream, world, rank, ph);
  if (dtype == GX_F32) return noise_draw_t<float>(slab, n, seed, GX_NOISE_FOURIER_PARALLEL, reverse_order, nullptr, (cudaStream_t) stream, world, rank, ph);
  return gx::fail("bad dtype");
}

// Parallel Fourier order on G ranks through the staged path: this rank generates its share of the shells (balanced by
// generator work) ONCE and stores every mode straight into the slab of the rank that owns its ky row -- peer_slabs[r] are
// peer-mapped pointers (gx_ipc_open) to the ranks' ky-slabs, peer_slabs[rank] == slab. With `spectrum` the factor
// sqrt(C(k)) / k^2 is applied as the modes are stored (the slab then holds delta_k / k^2, as after gx_slab_apply_transfer
// with divide_k2). The caller brackets the call with barriers across the ranks: nobody may still read the slabs when the
// first rank starts, and all ranks must have finished before anybody reads its slab.
int gx_slab_noise_staged(void *slab, int n, int dtype, int world, int rank, unsigned long seed, int reverse_order,
                         void *const *peer_slabs, const gx_spectrum *spectrum, double grid_length, void *scratch,
                         size_t scratch_bytes, void *stream) {
  if (!gx::is_pow2(n) || n < 4 || n > 4096) return gx::fail("grid size must be a power of two in [4,4096]");
  if (world < 1 || world > kMaxPeers || !gx::is_pow2(world) || n % world != 0 || rank < 0 || rank >= world)
    return gx::fail("gx_slab_noise_staged: bad world/rank (at most 8 ranks, powers of two)");
  Shard sh;
  sh.world = world; sh.rank = rank;
  for (int r = 0; r < world; ++r) {
    sh.peers[r] = r == rank ? slab : (peer_slabs ? peer_slabs[r] : nullptr);
    if (!sh.peers[r]) return gx::fail("gx_slab_noise_staged: null peer pointer");
  }
  cudaStream_t cs = (cudaStream_t) stream;
  const void *table = nullptr;
  if (spectrum)
    if (int rc = gx::build_transfer_table(n, dtype, grid_length, spectrum, 0.5, 1, cs, &table)) return rc;
  if (dtype == GX_F64) return noise_parallel_staged<double>((double *) slab, n, seed, reverse_order, (const double *) table, scratch, scratch_bytes, cs, sh);
  if (dtype == GX_F32) return noise_parallel_staged<float>((float *) slab, n, seed, reverse_order, (const float *) table, scratch, scratch_bytes, cs, sh);
  return gx::fail("bad dtype");
}

int gx_workspace_release(void) {
  gx::rng::ws_release();
  gx::release_spectral_workspace();
  gx::release_multigrid_workspace();
  gx::release_reduce_workspace();
  gx::pool_release();
  gx::fftgen::release_workspace();
  return 0;
}

int gx_rng_raw(unsigned long seed, unsigned long long offset, long n, uint32_t *host_out) {
  if (ensure_jump_table()) return 1;
  StreamDesc sd;
  sd.s0 = seed_state(seed); sd.xs = state_to_big(sd.s0); sd.b0 = 0; sd.entry0 = 0; sd.first_chunk = 0; sd.n_gauss = 0; sd.shell = 0;
  const long long q0 = (long long) (offset / kChunkLen);
  const long long q1 = (long long) ((offset + (unsigned long long) n + kChunkLen - 1) / kChunkLen);
  const int nch = (int) (q1 - q0);
  uint32_t *d = nullptr;
  GX_CUDA_CHECK(cudaMalloc(&d, sizeof(uint32_t) * (size_t) nch * kChunkLen));
  raw_chunks_kernel<<<(nch + 63) / 64, 64>>>(sd, g_jump, q0, nch, d);
  count_launch();
  GX_LAUNCH_CHECK();
  int rc = cudaMemcpy(host_out, d + (offset - (unsigned long long) q0 * kChunkLen), sizeof(uint32_t) * (size_t) n,
                      cudaMemcpyDeviceToHost) != cudaSuccess;
  cudaFree(d);
  return rc ? gx::fail("gx_rng_raw: copy failed") : 0;
}

int gx_rng_gaussians(unsigned long seed, long n, double *host_out, unsigned long long *raw_consumed) {
  double *d = nullptr;
  GX_CUDA_CHECK(cudaMalloc(&d, sizeof(double) * (size_t) n));
  Dest<double> dest;
  dest.mode = MODE_LIST; dest.n = 0; dest.reverse = 0; dest.field = nullptr; dest.list = d; dest.sigma = 1.0; dest.wtab = nullptr;
  dest.iy0 = 0; dest.ny = 0; dest.row0 = 0; dest.nrows = 0;
  std::vector<StreamReq> reqs;
  reqs.push_back({seed, raw_consumed ? *raw_consumed : 0ull, (long long) n, 0});
  std::vector<unsigned long long> ends;
  int rc = run_streams<double>(reqs, dest, &ends, 0);
  if (!rc) rc = cudaMemcpy(host_out, d, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost) != cudaSuccess;
  cudaFree(d);
  if (!rc && raw_consumed) *raw_consumed = ends[0];
  return rc;
}

}  // extern "C"

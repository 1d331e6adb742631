// Staged white noise: every chunk of a ranlxs2 stream is generated ONCE.
//
// The two-pass scheme of round 1 (rng.cu: pass A parses each chunk from its 4 possible entry positions, a scan resolves
// the true one, pass B regenerates the chunk and scatters every draw to its slot) ran the 397-step generator advance
// twice per chunk and wrote the field as isolated 16-byte stores. Here the chunk's accepted ziggurat draws are kept:
//   gen      one thread per chunk: jump, generate, parse the (up to) 4 entry hypotheses, and APPEND each accepted draw
//            as 32 bits -- (sign | level i) in the low byte, the 24-bit j above it; x = j * wtab[i] is re-formed exactly
//            on delivery -- to the list of the hypothesis it belongs to. Hypotheses merge after a few attempts (same
//            parity) or ~80 attempts (opposite parity), after which only list 0 grows: ~1.1 lists' worth per chunk.
//   scan     the chunk transfer functions (unchanged) give every chunk its true entry position and first gaussian index;
//   resolve  the true path of a chunk is  list_e[0, len1) ++ list_0[join, ...)  (chains through a second alt list are
//            spliced once), recorded as 16 bytes per chunk;
//   deliver  CALL-driven: a thread owns one (shell, k1, k2) position of drawRandomForSpecifiedGridFourier's loops
//            (randomfieldgenerator.hpp:226-243), i.e. three drawOneFourierMode calls = 6 consecutive gaussians = 24
//            contiguous staged bytes, and stores the three modes with the reference's overwrite rule resolved
//            analytically (a mode and its conjugate partner can both be drawn in a shell: the later call wins,
//            :144-171). Lanes run along k2, so the two faces with kz = k2 store whole contiguous row segments and the
//            face with kz = ks goes through a shared-memory transpose over 8 shells (128-byte row segments).
// Never-drawn slots (DC, the Nyquist planes and the shell edges (A, A, c <= A), (-A, -A, c < A)) are zeroed by a small
// kernel instead of a memset of the whole field.
//
// Everything here is __host__ __device__: tests/host/rng_staged_host.cu runs the same code on the CPU against the oracle.
#pragma once
#include "rng_core.cuh"
#include <type_traits>

namespace gx {
namespace rng {

constexpr int kBc = 64;               // increments per chunk
constexpr int kChunkLen = kBc * 24;   // raw outputs per chunk
constexpr int kCap = 772;             // list capacity: a chunk holds at most (1536 + 3) / 2 accepted draws
constexpr int kChunkStage = 4 * kCap; // 32-bit words of staging per chunk (4 hypothesis lists)
constexpr int kRing = 32;             // raw outputs kept per thread (the parse looks back < 28)

struct StreamDesc {
  State s0;              // seed state (used when the stream starts at block 0)
  Big xs;                // X at block b0 (start block of the stream's first chunk)
  long long b0;          // first block index of the stream's first chunk
  int entry0;            // position of the first attempt inside the first chunk
  int first_chunk;       // index of the stream's first chunk in the global chunk list
  long long n_gauss;     // gaussians wanted from this stream
  int shell;             // ks for per-shell streams, 0 otherwise
};

struct ChunkFn {
  long long count[4];
  unsigned char exit[4];
  unsigned char reset;
};

struct ScanState {
  long long base;
  int entry;
};

GX_HD ChunkFn compose(const ChunkFn &f, const ChunkFn &g) {  // f first, then g
  if (g.reset) return g;
  ChunkFn h;
  h.reset = f.reset;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h.exit[e] = g.exit[f.exit[e]];
    h.count[e] = f.count[e] + g.count[f.exit[e]];
  }
  return h;
}

GX_HD ScanState apply_fn(const ChunkFn &f, const ScanState &s) {
  ScanState o;
  if (f.reset) { o.entry = f.exit[0]; o.base = f.count[0]; }
  else { o.entry = f.exit[s.entry]; o.base = s.base + f.count[s.entry]; }
  return o;
}

// ---- gen -----------------------------------------------------------------------------------------------------------
struct ListMeta {          // how the 4 hypothesis lists of a chunk hang together
  unsigned short len[4];   // draws appended to list e before it merged into list rep[e] (list 0: all of its draws)
  unsigned short join[4];  // index in list rep[e] at which path e continues
  signed char rep[4];      // -1: never merged
};

struct StagedTail {        // an accepted tail draw seen by hypothesis `list` (value needs the host's log())
  int chunk;               // batch-relative chunk index
  unsigned int u1;
  unsigned short idx;      // position in that list
  unsigned char list, positive;
};

#ifdef __CUDA_ARCH__
#define GX_SYNCWARP() __syncwarp()
#else
#define GX_SYNCWARP()
#endif

GX_HD uint32_t pack_draw(uint32_t r0, uint32_t r1) { return (r0 & 0xFFu) | (r1 << 8); }

// x = j * wtab[i], value = sign * x: the products gsl_ran_gaussian_ziggurat forms (zig_attempt in rng_core.cuh)
GX_HD double unpack_draw(uint32_t e, const double *wtab) {
  const double x = GX_DMUL((double) (e >> 8), wtab[e & 0x7Fu]);
  const double sign = (e & 0x80u) ? 1.0 : -1.0;
  return GX_DMUL(GX_DMUL(sign, 1.0), x);
}

#ifdef __CUDA_ARCH__
#define GX_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define GX_ANY(x) (x)
#endif

// generator state of a chunk thread. Device: MODE 0 / 1 the (lo, hi) carry-flag form (mask on the alu / fma pipe), MODE 2 the
// FP64 form; host: the plain words
template <int MODE> struct GenState {
#ifdef __CUDA_ARCH__
  typename std::conditional<MODE == 2, StateF64, StateCC>::type s;
  GX_HD void load(const State &st) { s.load(st.x, st.c); }
  GX_HD void advance() {
    if constexpr (MODE == 2) s.advance();
    else s.template advance<MODE == 1>();
  }
  GX_HD uint32_t out(int j) const { return s.out(j); }
#else
  State s;
  GX_HD void load(const State &st) { s = st; }
  GX_HD void advance() { advance_block(s); }
  GX_HD uint32_t out(int j) const { return block_output(s, j); }
#endif
};

template <int E, class Buf> GX_HD void append_draw(uint32_t d, int c, Buf buf, uint32_t *stage) {
  if (E == 0) {  // the long list leaves as 16-byte stores
    buf.at(c & 3) = d;
    if ((c & 3) == 3) {
#ifdef __CUDA_ARCH__
      *reinterpret_cast<uint4 *>(stage + (c - 3)) = make_uint4(buf.at(0), buf.at(1), buf.at(2), d);
#else
      for (int k = 0; k < 4; ++k) stage[c - 3 + k] = buf.at(k);
#endif
    }
  } else {
    stage[E * kCap + c] = d;
  }
}

// Hypothesis E of a chunk parsed up to raw position `limit`. ALL lanes of the warp call this together (`live` says whether
// the lane has work). 98.8 % of the attempts are the integer fast path (j < ktab[i]); the others need exp() (wedge) or log()
// and exp() (tail), ~130 instructions. How far an attempt advances the parse does NOT depend on whether it is accepted
// (2 raw outputs on the fast path, 3 in a wedge, 4 in the tail), only the list of accepted draws does. So the lane walks
// its ~12 attempts of the increment in one loop of uniform length, collecting the draws in a small buffer `inc` with the
// slow attempts entered TENTATIVELY; then the pending slow attempts of all lanes are evaluated together (one convergent
// pass, two for a lane with two of them), rejected ones are taken out of the buffer, and the buffer is appended to the list.
// (Before: a lane that met a slow attempt waited for the others -- 24 loop iterations per increment at 15 of 32 lanes, ncu.)
constexpr int kIncCap = 16;   // draws of one increment per hypothesis: at most (24 + 3) / 2 = 13

template <int E, class Ring, class Buf, class Inc, class Sink>
GX_HD void parse_hypothesis(bool live, int &p, int &c, int limit, Ring ring, Buf buf, Inc inc, const unsigned int *ktab,
                            const double *wtab, const double *ytab, uint32_t *stage, Sink &sink) {
  bool more = live && p < limit;
  while (GX_ANY(more)) {           // one round, unless a lane meets a third slow attempt in the increment (~1e-4)
    int m = 0, npend = 0;          // draws collected in inc; pending slow attempts (position in inc, raw position)
    int pk0 = 0, pp0 = 0, pk1 = 0, pp1 = 0;
    if (more) {
      while (p < limit) {
        const uint32_t r0 = ring.at(p & (kRing - 1)), r1 = ring.at((p + 1) & (kRing - 1));
        const uint32_t i = r0 & 0x7Fu;
        if (r1 < ktab[i]) {
          inc.at(m++) = pack_draw(r0, r1);
          p += 2;
        } else if (npend < 2) {    // entered tentatively, decided after the loop
          inc.at(m) = pack_draw(r0, r1);
          if (npend == 0) { pk0 = m; pp0 = p; } else { pk1 = m; pp1 = p; }
          ++npend;
          ++m;
          p += (i < 127u) ? 3 : 4;
        } else {
          break;
        }
      }
    }
    GX_SYNCWARP();
    int removed = 0;               // a rejected attempt leaves the buffer: later entries move down by one
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
      const bool have = round < npend;
      if (!GX_ANY(have)) break;
      if (have) {
        const int pk = round == 0 ? pk0 : pk1, pp = round == 0 ? pp0 : pp1;
        const Attempt a = zig_attempt(ring.at(pp & (kRing - 1)), ring.at((pp + 1) & (kRing - 1)),
                                      ring.at((pp + 2) & (kRing - 1)), ring.at((pp + 3) & (kRing - 1)), ktab, wtab, ytab);
        const int k = pk - removed;
        if (a.accepted) {
          if (a.tail) sink.push(E, c + k, a.tail_u1, a.value > 0);   // entries before k are final: so is this index
        } else {
          for (int q = k; q + 1 < m; ++q) inc.at(q) = inc.at(q + 1);
          --m;
          ++removed;
        }
      }
    }
    for (int k = 0; k < m; ++k) { append_draw<E>(inc.at(k), c, buf, stage); ++c; }
    more = live && p < limit;
  }
}

// One chunk: `st` is the generator state at the chunk's first block. Ring::at(slot) / Buf::at(k) are the thread's raw-output
// ring (kRing slots), its 4-entry store buffer and its per-increment draw buffer (kIncCap slots; shared memory on the device,
// plain arrays on the host); `stage` is the
// chunk's kChunkStage words; sink.push(list, idx, u1, positive) records accepted tail draws. Warp-collective on the
// device: every lane of the warp must call it.
template <int MODE, class Ring, class Buf, class Inc, class Sink>
GX_HD void gen_chunk(const StreamDesc &sd, bool first, const State &st0, Ring ring, Buf buf, Inc inc, const unsigned int *ktab,
                     const double *wtab, const double *ytab, uint32_t *stage, ChunkFn &f, ListMeta &meta, Sink &sink) {
  int p0, p1 = 1, p2 = 2, p3 = 3;          // positions of the four entry hypotheses
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;      // draws appended to their lists
  int rep1 = -1, rep2 = -1, rep3 = -1;     // list a hypothesis merged into (a lower one), -1 while it is on its own
  int j1 = 0, j2 = 0, j3 = 0;              // index in that list at which its path continues
  p0 = first ? sd.entry0 : 0;
  if (first) rep1 = rep2 = rep3 = 0;       // a stream's first chunk starts at a known position: one hypothesis
  GenState<MODE> gs;
  gs.load(st0);

  for (int bi = 0; bi <= kBc; ++bi) {
    GX_SYNCWARP();
    gs.advance();
    const int rbase = 24 * bi;
#pragma unroll
    for (int j = 0; j < 24; ++j) ring.at((rbase + j) & (kRing - 1)) = gs.out(j);
    const int avail = 24 * (bi + 1);
    const int limit = avail - 3 < kChunkLen ? avail - 3 : kChunkLen;
    parse_hypothesis<0>(true, p0, c0, limit, ring, buf, inc, ktab, wtab, ytab, stage, sink);
    if (GX_ANY(rep1 < 0)) parse_hypothesis<1>(rep1 < 0, p1, c1, limit, ring, buf, inc, ktab, wtab, ytab, stage, sink);
    if (GX_ANY(rep2 < 0)) parse_hypothesis<2>(rep2 < 0, p2, c2, limit, ring, buf, inc, ktab, wtab, ytab, stage, sink);
    if (GX_ANY(rep3 < 0)) parse_hypothesis<3>(rep3 < 0, p3, c3, limit, ring, buf, inc, ktab, wtab, ytab, stage, sink);
    // merge hypotheses that have reached the same position (into the lowest one)
    if (rep1 < 0 && p1 == p0) { rep1 = 0; j1 = c0; }
    if (rep2 < 0) {
      if (p2 == p0) { rep2 = 0; j2 = c0; }
      else if (rep1 < 0 && p2 == p1) { rep2 = 1; j2 = c1; }
    }
    if (rep3 < 0) {
      if (p3 == p0) { rep3 = 0; j3 = c0; }
      else if (rep1 < 0 && p3 == p1) { rep3 = 1; j3 = c1; }
      else if (rep2 < 0 && p3 == p2) { rep3 = 2; j3 = c2; }
    }
  }
  for (int k = c0 & ~3; k < c0; ++k) stage[k] = buf.at(k & 3);

  const int pos[4] = {p0, p1, p2, p3}, cnt[4] = {c0, c1, c2, c3}, rep[4] = {-1, rep1, rep2, rep3}, joinv[4] = {0, j1, j2, j3};
  f.reset = first ? 1 : 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int r = first ? 0 : e;
    long long total = 0;
    int start = 0;
    for (int hop = 0; hop < 3 && rep[r] >= 0; ++hop) { total += cnt[r] - start; start = joinv[r]; r = rep[r]; }  // 3 -> 1 -> 0
    total += cnt[r] - start;
    f.exit[e] = (unsigned char) (pos[r] - kChunkLen);
    f.count[e] = total;
    meta.len[e] = (unsigned short) (first && e > 0 ? 0 : cnt[e]);
    meta.join[e] = (unsigned short) (first ? 0 : joinv[e]);
    meta.rep[e] = (signed char) (first && e > 0 ? 0 : rep[e]);
  }
}

// ---- resolve -------------------------------------------------------------------------------------------------------
struct ChunkLoc {
  long long base;            // stream-local index of the chunk's first gaussian
  unsigned short len1, join; // gaussian i of the chunk: i < len1 ? list_e[i] : list_0[i - len1 + join]
  unsigned char e, pad[3];
};

GX_HD ChunkLoc resolve_chunk(const ListMeta &m, int entry, long long base, uint32_t *stage) {
  ChunkLoc L;
  L.base = base; L.pad[0] = L.pad[1] = L.pad[2] = 0;
  const int e = entry;
  if (e == 0) { L.e = 0; L.len1 = 0; L.join = 0; return L; }
  int len = m.len[e], r = m.rep[e], j = m.join[e];
  while (r > 0) {  // the path runs through another alt list before it reaches list 0: splice that stretch onto ours
    const int lr = m.len[r];
    for (int k = j; k < lr; ++k) stage[e * kCap + len++] = stage[r * kCap + k];
    j = m.join[r];
    r = m.rep[r];
  }
  L.e = (unsigned char) e; L.len1 = (unsigned short) len; L.join = (unsigned short) (r == 0 ? j : 0);
  return L;
}

GX_HD uint32_t fetch_draw(const ChunkLoc &L, const uint32_t *stage, int i) {
  return i < (int) L.len1 ? stage[(int) L.e * kCap + i] : stage[i - (int) L.len1 + (int) L.join];
}

// index in the chunk's true path of the draw that hypothesis `list` stored at position idx, or -1 if it is not on the path
GX_HD int path_index(const ListMeta &m, int entry, int list, int idx) {
  int cur = entry, start = 0, off = 0;
  for (int hop = 0; hop < 4 && cur >= 0; ++hop) {
    if (cur == list) return (idx >= start && idx < (int) m.len[cur]) ? off + idx - start : -1;
    off += (int) m.len[cur] - start;
    start = m.join[cur];
    cur = m.rep[cur];
  }
  return -1;
}

// the chunk of a stream (chunks [fc, fc + nch) of `loc`) that holds gaussian g; `guess_scale` ~ chunks per gaussian of this
// stream ((nch - 1) / base of its last chunk): the counts per chunk scatter by a few draws, so the guess is off by a
// chunk or two at most and the probes below finish it
GX_HD int locate_chunk(const ChunkLoc *loc, int fc, int nch, long long g, float guess_scale) {
  int k = (int) ((float) g * guess_scale);
  if (k > nch - 1) k = nch - 1;
  while (k > 0 && loc[fc + k].base > g) --k;
  while (k + 1 < nch && loc[fc + k + 1].base <= g) ++k;
  return fc + k;
}

GX_HD float locate_scale(const ChunkLoc *loc, int fc, int nch) {
  return nch > 1 ? (float) ((double) (nch - 1) / (double) loc[fc + nch - 1].base) : 0.0f;
}

// ---- deliver -------------------------------------------------------------------------------------------------------
// drawOneFourierMode call `face` at loop position (ks; p, q): the mode (k1, k2, k3), whether this call is the last writer
// of its slot, and where setFourierCoefficient stores it.
struct FaceSlot {
  int a, b, c;       // signed coordinates of the stored representative (c >= 0)
  bool neg;          // the representative is -k: the imaginary part is stored negated
  bool write;        // this call is the last writer (its conjugate partner is not drawn later in the shell)
  bool second;       // k3 == 0: drawOneFourierMode's explicit second write lands on a different slot, (-k1, -k2, 0)
  int m;             // k1^2 + k2^2 + k3^2
};

GX_HD FaceSlot face_slot(int ks, int p, int q, int face) {
  const int k1 = (face == 0) ? ks : p;
  const int k2 = (face == 0) ? p : ((face == 1) ? ks : q);
  const int k3 = (face == 2) ? ks : q;
  if (p != -ks && q != -ks) {
    // no coordinate equals -ks: the conjugate partner has no coordinate equal to +ks, so it is not drawn in this shell and
    // this call is the only writer of its slot (all but one lane per row and one row per shell take this path)
    FaceSlot s;
    s.write = true;
    s.m = k1 * k1 + k2 * k2 + k3 * k3;
    s.neg = k3 < 0;
    s.a = s.neg ? -k1 : k1; s.b = s.neg ? -k2 : k2; s.c = s.neg ? -k3 : k3;
    s.second = (k3 == 0);
    return s;
  }
  // the partner -k is drawn in the same shell iff exactly one of its coordinates equals +ks and the others lie in [-ks, ks-1]
  const int a = -k1, b = -k2, c3 = -k3;
  const int np = (a == ks) + (b == ks) + (c3 == ks);
  const int pface = (a == ks) ? 0 : ((b == ks) ? 1 : 2);
  const int pp = (pface == 0) ? b : a;
  const int pq = (pface == 2) ? b : c3;
  const bool partner_drawn = (np == 1) && (pp >= -ks) && (pp < ks) && (pq >= -ks) && (pq < ks);
  const int twoks = 2 * ks;
  const int c_self = ((p + ks) * twoks + (q + ks)) * 3 + face;
  const int c_partner = ((pp + ks) * twoks + (pq + ks)) * 3 + pface;
  FaceSlot s;
  s.write = !(partner_drawn && c_partner > c_self);
  s.m = k1 * k1 + k2 * k2 + k3 * k3;
  s.neg = k3 < 0;
  s.a = s.neg ? -k1 : k1; s.b = s.neg ? -k2 : k2; s.c = s.neg ? -k3 : k3;
  s.second = (k3 == 0);
  return s;
}

// the two draws of a call in the field's type: norm * gaussian with norm = FloatType(1/sqrt(2)) (randomfieldgenerator.hpp:213),
// times the optional per-mode factor (Field::applyTransferFunction folded in), re/im swapped for the reverse order
template <typename T>
GX_HD void call_value(double first, double second, double sigma, const T *wtab, int m, int reverse, T &re, T &im) {
  T a = (T) GX_DMUL(sigma, first), b = (T) GX_DMUL(sigma, second);
  if (wtab) {
#ifdef __CUDA_ARCH__
    const T w = __ldg(wtab + m);
#else
    const T w = wtab[m];
#endif
    a = (T) GX_DMUL(a, w); b = (T) GX_DMUL(b, w);
  }
  re = reverse ? b : a;
  im = reverse ? a : b;
}

// slots of the parallel Fourier order that no call ever writes (they keep the zero of the freshly constructed field)
GX_HD bool never_drawn(int n, int ix, int iy, int iz) {
  const int h = n >> 1;
  if (ix == h || iy == h || iz == h) return true;
  const int a = ix <= h ? ix : ix - n, b = iy <= h ? iy : iy - n, c = iz;
  const int aa = a < 0 ? -a : a, ab = b < 0 ? -b : b;
  int M = aa > ab ? aa : ab;
  if (c > M) M = c;
  if (M == 0) return true;  // DC
  // k is drawn (in shell M) iff exactly one coordinate equals +M; -k iff exactly one coordinate of k equals -M
  const int plus = (a == M) + (b == M) + (c == M), minus = (a == -M) + (b == -M);
  return plus != 1 && minus != 1;
}

}  // namespace rng
}  // namespace gx

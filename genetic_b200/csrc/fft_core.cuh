// In-register / shared-memory 1-D complex FFT building block (power-of-two lengths 2..4096).
//
// Replaces the FFTW plans of the reference (tools/numerics/fourier.hpp:521-563,
// fftw_plan_dft_r2c_3d / c2r_3d) as the line transform inside each of the three 1-D passes.
//
// Structure: mixed-radix decimation-in-frequency, radices from {2,4,8,16}. A line of NC points is
// owned by P = NC/E threads, each holding E points in registers (E = first radix). Stage 0 works on
// the registers as loaded (thread t holds x[t + i*P]); later stages exchange through one shared-memory
// copy of the line (in place: a butterfly reads and writes the same R positions), with one
// __syncthreads between stages. The last stage hands each output to an `emit(k, value)` callback with
// its natural frequency index k (mixed-radix digit reversal is resolved in index arithmetic, not by a
// data pass). Everything here is __host__ __device__ so the index math is unit-tested on the CPU
// (tests/host/fft_core_host_test.cu) -- there is no GPU in the build container.
#pragma once
#include "gx_common.cuh"

namespace gx {
namespace fft {

// DIR = -1: forward, kernel exp(-2 pi i nk/N) (FFTW_FORWARD); DIR = +1: backward.
template <int DIR, typename C> GX_HD C mul_i_dir(C a) {  // a * (DIR * i)
  C r;
  if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
  return r;
}

template <typename T> struct K {
  static constexpr T rs2 = (T) 0.70710678118654752440084436210485L;  // 1/sqrt(2)
  static constexpr T c16 = (T) 0.92387953251128675612818318939679L;  // cos(pi/8)
  static constexpr T s16 = (T) 0.38268343236508977172845998403040L;  // sin(pi/8)
};

// a * W8^k, W8 = exp(DIR * 2 pi i / 8)
template <int DIR, int k, typename T> GX_HD cplx<T> mul_w8(cplx<T> a) {
  constexpr T h = K<T>::rs2;
  cplx<T> r;
  if (k == 0) return a;
  if (k == 2) return mul_i_dir<DIR>(a);
  if (k == 1) {
    if (DIR < 0) { r.x = (a.x + a.y) * h; r.y = (a.y - a.x) * h; }
    else { r.x = (a.x - a.y) * h; r.y = (a.x + a.y) * h; }
    return r;
  }
  // k == 3
  if (DIR < 0) { r.x = (a.y - a.x) * h; r.y = -(a.x + a.y) * h; }
  else { r.x = -(a.x + a.y) * h; r.y = (a.x - a.y) * h; }
  return r;
}

// a * W16^k for k in {1,3,9} (the others reduce to W8 powers)
template <int DIR, int k, typename T> GX_HD cplx<T> mul_w16(cplx<T> a) {
  constexpr T c = K<T>::c16, s = K<T>::s16;
  // W16^1 = (c, DIR*s); W16^3 = (s, DIR*c); W16^9 = -(W16^1)
  T wr = (k == 3) ? s : c, wi = (k == 3) ? c : s;
  if (DIR < 0) wi = -wi;
  cplx<T> r;
  r.x = a.x * wr - a.y * wi;
  r.y = a.x * wi + a.y * wr;
  if (k == 9) { r.x = -r.x; r.y = -r.y; }
  return r;
}

template <int DIR, typename C> GX_HD void r2(C &a0, C &a1) {
  C t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

// natural-order radix-4: (a0..a3) <- DFT4
template <int DIR, typename C> GX_HD void r4(C &a0, C &a1, C &a2, C &a3) {
  C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i_dir<DIR>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}

template <int DIR, typename T> GX_HD void r8(cplx<T> *a) {
  using C = cplx<T>;
  r4<DIR>(a[0], a[2], a[4], a[6]);  // E0..E3 in a0,a2,a4,a6
  r4<DIR>(a[1], a[3], a[5], a[7]);  // O0..O3 in a1,a3,a5,a7
  C o0 = a[1], o1 = mul_w8<DIR, 1, T>(a[3]), o2 = mul_w8<DIR, 2, T>(a[5]), o3 = mul_w8<DIR, 3, T>(a[7]);
  C e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
  a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
  a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
  a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
  a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <int DIR, typename T> GX_HD void r16(cplx<T> *a) {
  using C = cplx<T>;
  // four DFT4 over a[j + 4r] (r = 0..3): a[j + 4q] <- B_j[q]
#pragma unroll
  for (int j = 0; j < 4; ++j) r4<DIR>(a[j], a[j + 4], a[j + 8], a[j + 12]);
  // twiddle B_j[q] *= W16^(j q)
  a[5] = mul_w16<DIR, 1, T>(a[5]);    // j=1,q=1
  a[9] = mul_w8<DIR, 1, T>(a[9]);     // j=1,q=2 -> W16^2
  a[13] = mul_w16<DIR, 3, T>(a[13]);  // j=1,q=3
  a[6] = mul_w8<DIR, 1, T>(a[6]);     // j=2,q=1 -> W16^2
  a[10] = mul_w8<DIR, 2, T>(a[10]);   // j=2,q=2 -> W16^4
  a[14] = mul_w8<DIR, 3, T>(a[14]);   // j=2,q=3 -> W16^6
  a[7] = mul_w16<DIR, 3, T>(a[7]);    // j=3,q=1
  a[11] = mul_w8<DIR, 3, T>(a[11]);   // j=3,q=2 -> W16^6
  a[15] = mul_w16<DIR, 9, T>(a[15]);  // j=3,q=3
  // DFT4 across j for each q: a[4q + p] <- y[q + 4p]
#pragma unroll
  for (int q = 0; q < 4; ++q) r4<DIR>(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
  // transpose to natural order: out[q + 4p] = a[4q + p]
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int p = q + 1; p < 4; ++p) { C t = a[4 * q + p]; a[4 * q + p] = a[4 * p + q]; a[4 * p + q] = t; }
}

template <int R, int DIR, typename T> GX_HD void butterfly(cplx<T> *a) {
  if (R == 2) r2<DIR>(a[0], a[1]);
  else if (R == 4) r4<DIR>(a[0], a[1], a[2], a[3]);
  else if (R == 8) r8<DIR, T>(a);
  else r16<DIR, T>(a);
}

// ---- per-length plan: radices (largest first), E = radices[0]
// SEL = 0: fewest stages (16 points per thread where possible). SEL = 1: 8 points per thread -- half the registers and
// twice the threads per line, for kernels that need more warps in flight (the f64 column passes).
template <int NC, int SEL = 0> struct Plan : Plan<NC, 0> {};  // SEL != 0 without a specialisation: the default plan
#define GX_PLAN(N_, S_, R0, R1, R2)                                  \
  template <> struct Plan<N_, 0> {                                   \
    static constexpr int S = S_;                                     \
    GX_HD static constexpr int radix(int s) { return s == 0 ? R0 : (s == 1 ? R1 : R2); } \
  };
GX_PLAN(2, 1, 2, 1, 1)
GX_PLAN(4, 1, 4, 1, 1)
GX_PLAN(8, 1, 8, 1, 1)
GX_PLAN(16, 1, 16, 1, 1)
GX_PLAN(32, 2, 8, 4, 1)
GX_PLAN(64, 2, 8, 8, 1)
GX_PLAN(128, 2, 16, 8, 1)
GX_PLAN(256, 2, 16, 16, 1)
GX_PLAN(512, 3, 8, 8, 8)
GX_PLAN(1024, 3, 16, 8, 8)
GX_PLAN(2048, 3, 16, 16, 8)
GX_PLAN(4096, 3, 16, 16, 16)
#undef GX_PLAN
template <> struct Plan<256, 1> {
  static constexpr int S = 3;
  GX_HD static constexpr int radix(int s) { return s == 0 ? 8 : (s == 1 ? 8 : 4); }
};


// shared-memory index skew: one pad element every SKEW elements keeps power-of-two strides off the
// same banks (SKEW elements == 128 bytes == one full bank row)
template <typename T> GX_HD constexpr int skew_of() { return 128 / (int) sizeof(cplx<T>); }
template <typename T> GX_HD int sidx(int p) { return p + p / skew_of<T>(); }
template <typename T> GX_HD constexpr int line_pitch(int nc) { return nc + nc / skew_of<T>() + 1; }

// LS = 0: the line is contiguous in shared memory with the bank skew above (row passes).
// LS > 0: element p of the line lives at line[p * LS] (column passes: LS adjacent columns are interleaved,
//         so a warp touches LS * sizeof(C) contiguous bytes per line position and no skew is needed).
template <typename T, int NC, int DIR, int LS = 0, int SEL = 0> struct BlockFFT {
  using C = cplx<T>;
  using PL = Plan<NC, (SEL & 1)>;  // SEL bit 0: alternative plan; bit 1: twiddle powers by multiplication in column passes
  static constexpr int S = PL::S;
  static constexpr int E = PL::radix(0);
  static constexpr int P = NC / E;

  GX_HD static int lidx(int p) { return LS > 0 ? p * LS : sidx<T>(p); }

  GX_HD static constexpr int L(int s) {  // sub-transform length entering stage s
    int l = NC;
    for (int i = 0; i < s; ++i) l /= PL::radix(i);
    return l;
  }
  GX_HD static constexpr int D(int s) { return L(s) / PL::radix(s); }

  // natural frequency index of final in-place position p
  GX_HD static int freq_of_pos(int p) {
    int k = 0, mult = 1;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int d = D(s), r = PL::radix(s);
      int q = (p / d) % r;
      k += q * mult;
      mult *= r;
    }
    return k;
  }

  // Twiddles of one butterfly are the powers w, w^2, .., w^(R-1) of one table entry. In double precision they are
  // formed by multiplication (w^(h+l) = w^h * w^l with h in {4, 8, 12}, l in {1, 2, 3}: depth <= 5, error a few 1e-16)
  // instead of R-1 table loads at stride q: those loads hit L1 but each touches up to 32 distinct 128-byte lines per warp,
  // and the L1 data pipe, not the FP64 pipe, is what the passes run out of (ncu: L1 data pipe 83 %, FP64 26 % in the z-pass).
  // (Row passes only, LS == 0: the column passes read their twiddles from a private shared-memory copy, their L1 data pipe
  // is at ~50 %, and the extra live values made them spill under the 96-register cap.)
  static constexpr bool kPowersByMultiplication = (sizeof(T) == 8) && (LS == 0 || (SEL & 2) != 0);
  // a[q] *= w1^q for q = 1..R-1 with few values live (w1..w3 and w4, w8, w12): w^(h+l) = w^h * w^l
  template <int R> GX_HD static void apply_twiddle_powers(C w1, C *a) {
    a[1] = cmul(a[1], w1);
    if (R == 2) return;
    const C w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    a[2] = cmul(a[2], w2);
    a[3] = cmul(a[3], w3);
    if (R == 4) return;
    C wh = cmul(w2, w2);  // w^4
#pragma unroll
    for (int h = 4; h < R; h += 4) {
      a[h] = cmul(a[h], wh);
      a[h + 1] = cmul(a[h + 1], cmul(wh, w1));
      a[h + 2] = cmul(a[h + 2], cmul(wh, w2));
      a[h + 3] = cmul(a[h + 3], cmul(wh, w3));
      if (h + 4 < R) wh = (h == 4) ? cmul(wh, wh) : cmul(wh, cmul(w2, w2));  // w^8 = (w^4)^2, w^12 = w^8 * w^4
    }
  }

  GX_HD static C twiddle(const C *__restrict__ tw, int j) {
    // tw[j] = exp(-2 pi i j / NC); the table may live in global memory (const __restrict__ -> LDG.CONSTANT) or in
    // shared memory (column passes: the cluster barrier's acquire invalidates L1, so they keep a private copy)
    C w = tw[j];
    if (DIR > 0) w.y = -w.y;
    return w;
  }

  // stage 0 on the registers as loaded: v[i] = x[t + i*P]
  template <class Emit> GX_HD static void stage0(C (&v)[E], int t, C *line, const C *__restrict__ tw, Emit &&emit) {
    butterfly<E, DIR, T>(v);
    if (S == 1) {
#pragma unroll
      for (int q = 0; q < E; ++q) emit(q, v[q]);  // NC == E, P == 1, t == 0
      return;
    }
    if (kPowersByMultiplication) {
      apply_twiddle_powers<E>(twiddle(tw, t), v);
    } else {
#pragma unroll
      for (int q = 1; q < E; ++q) v[q] = cmul(v[q], twiddle(tw, t * q));
    }
#pragma unroll
    for (int q = 0; q < E; ++q) line[lidx(t + q * P)] = v[q];
  }

  template <int s, class Emit> GX_HD static void stage(int t, C *line, const C *__restrict__ tw, Emit &&emit) {
    constexpr int R = PL::radix(s);
    constexpr int Ls = L(s), Ds = D(s);
    constexpr bool last = (s == S - 1);
#pragma unroll
    for (int j = 0; j < E / R; ++j) {
      const int b = t + j * P;
      const int g = b / Ds, m = b % Ds;
      const int base = g * Ls + m;
      C a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = line[lidx(base + r * Ds)];
      butterfly<R, DIR, T>(a);
      if (last) {
#pragma unroll
        for (int q = 0; q < R; ++q) emit(freq_of_pos(base + q * Ds), a[q]);
      } else {
        if (kPowersByMultiplication) {
          apply_twiddle_powers<R>(twiddle(tw, m * (NC / Ls)), a);
        } else {
#pragma unroll
          for (int q = 1; q < R; ++q) a[q] = cmul(a[q], twiddle(tw, m * q * (NC / Ls)));
        }
#pragma unroll
        for (int q = 0; q < R; ++q) line[lidx(base + q * Ds)] = a[q];
      }
    }
  }

#ifdef __CUDACC__
  // Last stage with every output held in registers until the whole CTA has finished reading the line, so that
  // emit() may overwrite the line buffer itself (used by the r2c row pass, which needs the spectrum in natural
  // order in shared memory and has no second buffer).
  template <int s, class Emit> __device__ __forceinline__ static void stage_last_synced(int t, C *line, Emit &&emit) {
    constexpr int R = PL::radix(s);
    constexpr int Ls = L(s), Ds = D(s);
    C out[E];
#pragma unroll
    for (int j = 0; j < E / R; ++j) {
      const int b = t + j * P;
      const int base = (b / Ds) * Ls + (b % Ds);
      C a[R];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = line[lidx(base + r * Ds)];
      butterfly<R, DIR, T>(a);
#pragma unroll
      for (int q = 0; q < R; ++q) out[j * R + q] = a[q];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < E / R; ++j) {
      const int b = t + j * P;
      const int base = (b / Ds) * Ls + (b % Ds);
#pragma unroll
      for (int q = 0; q < R; ++q) emit(freq_of_pos(base + q * Ds), out[j * R + q]);
    }
  }

  template <class Emit> __device__ __forceinline__ static void run_synced_emit(C (&v)[E], int t, C *line,
                                                                              const C *__restrict__ tw, Emit &&emit) {
    if constexpr (S == 1) {
      stage0(v, t, line, tw, emit);
    } else if constexpr (S == 2) {
      stage0(v, t, line, tw, emit);
      __syncthreads();
      stage_last_synced<1>(t, line, emit);
    } else {
      stage0(v, t, line, tw, emit);
      __syncthreads();
      stage<1>(t, line, tw, emit);
      __syncthreads();
      stage_last_synced<2>(t, line, emit);
    }
  }

  // Device driver. All threads of the CTA must call it (it contains __syncthreads).
  template <class Emit> __device__ __forceinline__ static void run(C (&v)[E], int t, C *line, const C *__restrict__ tw,
                                                                   Emit &&emit) {
    stage0(v, t, line, tw, emit);
    if constexpr (S >= 2) {
      __syncthreads();
      stage<1>(t, line, tw, emit);
    }
    if constexpr (S >= 3) {
      __syncthreads();
      stage<2>(t, line, tw, emit);
    }
  }
#endif
};

// ---- row transform with its outputs in NATURAL order ----------------------------------------------------------------------
// BlockFFT is decimation in frequency: its last stage leaves frequency freq_of_pos(p) at position p, so the lanes of a warp
// emit digit-reversed indices and the c2r pass, which stores straight from that stage, issues 8..16-byte stores to 32
// different sectors per instruction (ncu, z-pass at 1024^3: L1 data pipe 89 %, more than half of it those stores). Here the
// LAST stage is decimation in time instead:
//     X[k1 + M' u] = sum_r W_16^(r u) * ( W_M^(r k1) * Y_r[k1] ),   M' = M / 16,  u = 0..15,
// where Y_r is the M'-point transform of the inputs n = r (mod 16). Thread t owns k1 = t, so for every u the lanes of a warp
// emit consecutive k: 256 contiguous bytes per store instruction. The 16 sub-transforms are decimation in frequency in one
// or two stages (radices RA, RB; RB == 1: one stage), in place in shared memory:
//     stage A   butterfly (r, j):  a_q = sum_i x[r + 16 (j + RB i)] W_RA^(i q),  a_q *= W_M'^(j q)   -> S[r][q + RA j]
//     stage B   butterfly (r, q):  Y_r[q + RA p] = sum_j S[r][q + RA j] W_RB^(j p)                    -> S[r][q + RA p]
//     final     thread k1: the formula above, reading S[r][k1] for r = 0..15.
// S[r][.] has pitch M' + 1 elements, which keeps the 16 sub-transforms of a butterfly column on different banks.
template <typename T, int M, int DIR> struct DitRows {
  using C = cplx<T>;
  static_assert(M >= 256 && M <= 2048, "DitRows: 256 <= M <= 2048");
  static constexpr int MP = M / 16;                                  // length of a sub-transform, threads per row
  static constexpr int RA = (MP == 16 || MP == 128) ? 16 : 8;
  static constexpr int RB = MP / RA;                                 // 1, 4, 8, 8
  static constexpr int P = MP;
  static constexpr int E = 16;
  static constexpr int NA = 16 / RA;                                 // stage-A butterflies per thread
  static constexpr int NB = (RB > 1) ? 16 / RB : 0;                  // stage-B butterflies per thread
  static constexpr int PITCH = MP + 1;
  static constexpr int smem_elems = 16 * PITCH;

  // input n (0 <= n < M) that thread t holds as element e = c * RA + i of its 16 (stage-A butterfly c, leg i)
  GX_HD static int input_index(int t, int e) {
    const int c = e / RA, i = e % RA;
    const int b = t + P * c, r = b % 16, j = b / 16;
    return r + 16 * (j + RB * i);
  }

  GX_HD static C tw_at(const C *__restrict__ twM, int idx) {          // W_M^idx, conjugated for the inverse
    C w = twM[idx];
    if (DIR > 0) w.y = -w.y;
    return w;
  }

  // v[e]: the 16 inputs of thread t in input_index order
  GX_HD static void stage_a(C (&v)[16], int t, C *S, const C *__restrict__ twM) {
#pragma unroll
    for (int c = 0; c < NA; ++c) {
      const int b = t + P * c, r = b % 16, j = b / 16;
      C *a = v + c * RA;
      butterfly<RA, DIR, T>(a);
      if (RB > 1) {
        // W_M'^(j q) = W_M^(16 j q)
        if (sizeof(T) == 8) BlockFFT<T, M, DIR>::template apply_twiddle_powers<RA>(tw_at(twM, 16 * j), a);
        else {
#pragma unroll
          for (int q = 1; q < RA; ++q) a[q] = cmul(a[q], tw_at(twM, 16 * j * q));
        }
      }
#pragma unroll
      for (int q = 0; q < RA; ++q) S[r * PITCH + q + RA * j] = a[q];
    }
  }

  GX_HD static void stage_b(int t, C *S) {
    if (RB == 1) return;
#pragma unroll
    for (int c = 0; c < (NB > 0 ? NB : 1); ++c) {
      const int b = t + P * c, q = b % RA, r = b / RA;
      C a[RB > 1 ? RB : 2];
#pragma unroll
      for (int j = 0; j < RB; ++j) a[j] = S[r * PITCH + q + RA * j];
      butterfly<RB, DIR, T>(a);
#pragma unroll
      for (int p2 = 0; p2 < RB; ++p2) S[r * PITCH + q + RA * p2] = a[p2];
    }
  }

  template <class Emit> GX_HD static void stage_final(int t, const C *S, const C *__restrict__ twM, Emit &&emit) {
    C y[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) y[r] = S[r * PITCH + t];
    if (sizeof(T) == 8) BlockFFT<T, M, DIR>::template apply_twiddle_powers<16>(tw_at(twM, t), y);
    else {
#pragma unroll
      for (int r = 1; r < 16; ++r) y[r] = cmul(y[r], tw_at(twM, t * r));
    }
    butterfly<16, DIR, T>(y);
#pragma unroll
    for (int u = 0; u < 16; ++u) emit(t + MP * u, y[u]);
  }
};

}  // namespace fft
}  // namespace gx

// gsl_rng_ranlxs2 + gsl_ran_gaussian_ziggurat, draw-for-draw, in a form that parallelises.
//
// The reference draws its white noise from GSL (randomfieldgenerator.hpp:43-45,149-150,191,228-229); GSL is an
// un-vendored dependency, so the algorithms are restated from their publications:
//   * ranlxs2 = Luescher's RANLUX (v3): subtract-with-borrow x_n = x_{n-5} - x_{n-12} - c (mod b), b = 2^48, on a
//     ring of 12 words; every "increment" advances 397 word steps and then releases the 12 words as 24 outputs
//     of 24 bits (low half, then high half of each word, oldest word first).
//   * The SWB recurrence is the lagged form of the linear congruential generator X_{n+1} = b^-1 X_n (mod m),
//     m = b^12 - b^5 + 1, with X_n = sum_{i<12} x_{n-12+i} b^i - sum_{i<5} x_{n-5+i} b^i + c_{n-1}
//     (Tezuka/L'Ecuyer/Couture 1993). That gives O(log k) skip-ahead: X_{n+k} = a^k X_n, a = b^-1 = m - b^11 + b^4,
//     and the state is read back from X by 12 "multiply by b, take the quotient" steps.
//   * ziggurat: 128 levels, two 24-bit outputs per attempt (+1 in the wedges, +2 in the tail) -- the number of
//     outputs a gaussian consumes is data dependent, so gaussian index -> output index is resolved by a scan
//     over per-chunk transfer functions (rng.cu).
// All of it is integer / IEEE-exact except log() and exp() in the rare tail and wedge branches (see rng.cu).
// __host__ __device__ throughout: unit-tested on the CPU against the oracle (tests/host/rng_core_host_test.cu).
#pragma once
#include "gx_common.cuh"

namespace gx {
namespace rng {

constexpr int kLimbs = 24;  // 576 bits as 24 x 24-bit limbs
constexpr uint32_t kMask24 = 0xFFFFFFu;
constexpr uint64_t kMask48 = 0xFFFFFFFFFFFFull;
constexpr int kStepsPerBlock = 397;  // ranlxs2 luxury
constexpr int kRawPerBlock = 24;

struct Big {  // value in [0, m), little-endian 24-bit limbs
  uint32_t l[kLimbs];
};

// m = 2^576 - 2^240 + 1
GX_HD uint32_t m_limb(int i) { return i == 0 ? 1u : (i >= 10 ? kMask24 : 0u); }

// Normalise signed columns (value = sum c[i] 2^(24 i), congruent to what we want mod m, value >= 0 overall)
// into [0, m). Returns the number of times m was subtracted (needed by the digit-extraction step).
GX_HD long long normalize(long long *c, Big &out) {
  long long subtracted = 0;
  for (int round = 0; round < 4; ++round) {
    long long carry = 0;
    for (int i = 0; i < kLimbs; ++i) {
      long long v = c[i] + carry;
      carry = v >> 24;  // arithmetic shift: floor division
      c[i] = v - (carry << 24);
    }
    if (carry == 0) break;
    // carry * 2^576 == carry * (m + 2^240 - 1)
    subtracted += carry;
    c[10] += carry;
    c[0] -= carry;
  }
  // now 0 <= value < 2^576; one conditional subtraction of m
  bool ge = true;
  for (int i = kLimbs - 1; i >= 0; --i) {
    const long long mi = (long long) m_limb(i);
    if (c[i] != mi) { ge = c[i] > mi; break; }
  }
  if (ge) {
    long long borrow = 0;
    for (int i = 0; i < kLimbs; ++i) {
      long long v = c[i] - (long long) m_limb(i) - borrow;
      borrow = v < 0;
      c[i] = v + (borrow << 24);
    }
    subtracted += 1;
  }
  for (int i = 0; i < kLimbs; ++i) out.l[i] = (uint32_t) c[i];
  return subtracted;
}

GX_HD Big modmul(const Big &a, const Big &b) {
  unsigned long long col[2 * kLimbs];
  for (int k = 0; k < 2 * kLimbs; ++k) col[k] = 0;
  for (int i = 0; i < kLimbs; ++i) {
    const unsigned long long ai = a.l[i];
    for (int j = 0; j < kLimbs; ++j) col[i + j] += ai * (unsigned long long) b.l[j];  // < 24 * 2^48 < 2^53
  }
  long long c[2 * kLimbs];
  for (int k = 0; k < 2 * kLimbs; ++k) c[k] = (long long) col[k];
  // fold 2^(24k), k >= 24:  2^576 == 2^240 - 1  =>  limb k -> +limb (k-14), -limb (k-24)
  for (int k = 2 * kLimbs - 1; k >= kLimbs; --k) {
    c[k - 14] += c[k];
    c[k - 24] -= c[k];
    c[k] = 0;
  }
  Big r;
  normalize(c, r);
  return r;
}

GX_HD Big big_one() { Big r; for (int i = 0; i < kLimbs; ++i) r.l[i] = 0; r.l[0] = 1; return r; }

// a = b^-1 mod m = m - b^11 + b^4   (b * (b^11 - b^4) = m - 1)
GX_HD Big big_a() {
  long long c[2 * kLimbs];
  for (int i = 0; i < 2 * kLimbs; ++i) c[i] = i < kLimbs ? (long long) m_limb(i) : 0;
  c[22] -= 1;  // b^11 = 2^528 = limb 22
  c[8] += 1;   // b^4 = 2^192 = limb 8
  Big r;
  normalize(c, r);
  return r;
}

GX_HD Big modpow(Big base, unsigned long long e) {
  Big r = big_one();
  while (e) {
    if (e & 1ull) r = modmul(r, base);
    base = modmul(base, base);
    e >>= 1;
  }
  return r;
}

// Generator state in time order: x[0] is the oldest word (x_{n-12}), x[11] the newest; c the pending borrow.
struct State {
  uint64_t x[12];
  uint32_t c;
};

// gsl_rng_set(ranlxs2, s): 31-bit seed -> 12 words of 48 bits through a 31-bit LFSR (bit k of the seed, taps 0/18)
GX_HD State seed_state(unsigned long s) {
  if (s == 0) s = 1;
  uint32_t v = (uint32_t) (s & 0x7fffffffUL);
  int xbit[31];
  for (int k = 0; k < 31; ++k) { xbit[k] = (int) (v & 1u); v >>= 1; }
  int ibit = 0, jbit = 18;
  State st;
  for (int k = 0; k < 12; ++k) {
    uint64_t x = 0;
    for (int mm = 0; mm < 48; ++mm) {
      x = (x << 1) | (uint64_t) xbit[ibit];
      xbit[ibit] = (xbit[ibit] + xbit[jbit]) & 1;
      ibit = (ibit + 1) % 31;
      jbit = (jbit + 1) % 31;
    }
    st.x[k] = x;  // ir = 0: word 0 is the first to be replaced, i.e. the oldest
  }
  st.c = 0;
  return st;
}

// X = sum_{i<12} x_i b^i - sum_{i<5} x_{7+i} b^i + c   (mod m)
GX_HD Big state_to_big(const State &st) {
  long long c[2 * kLimbs];
  for (int i = 0; i < 2 * kLimbs; ++i) c[i] = 0;
  for (int i = 0; i < 12; ++i) {
    c[2 * i] += (long long) (st.x[i] & kMask24);
    c[2 * i + 1] += (long long) (st.x[i] >> 24);
  }
  for (int i = 0; i < 5; ++i) {
    c[2 * i] -= (long long) (st.x[7 + i] & kMask24);
    c[2 * i + 1] -= (long long) (st.x[7 + i] >> 24);
  }
  c[0] += st.c;
  // the sum may be negative (> -b^5): add m once so that normalize sees a non-negative value
  for (int i = 0; i < kLimbs; ++i) c[i] += (long long) m_limb(i);
  Big r;
  normalize(c, r);
  return r;
}

// Inverse of state_to_big: Y <- b*Y mod m twelve times; the quotients are the words, newest first.
GX_HD State big_to_state(const Big &X) {
  State st;
  Big Y = X;
  for (int d = 0; d < 12; ++d) {
    const unsigned long long ytop = (unsigned long long) Y.l[22] | ((unsigned long long) Y.l[23] << 24);
    long long c[2 * kLimbs];
    for (int i = 0; i < 2 * kLimbs; ++i) c[i] = 0;
    for (int i = 0; i < 22; ++i) c[i + 2] = (long long) Y.l[i];  // (Y mod b^11) * b
    // ytop * 2^576 == ytop * m + ytop * (2^240 - 1)
    c[10] += (long long) (ytop & kMask24);
    c[11] += (long long) (ytop >> 24);
    c[0] -= (long long) (ytop & kMask24);
    c[1] -= (long long) (ytop >> 24);
    // keep it non-negative for normalize (ytop*(2^240-1) >= ytop, so the total already is; the borrow
    // out of limbs 0/1 is repaid from limb 2 upward by the signed carry propagation)
    const long long extra = normalize(c, Y);
    st.x[11 - d] = (uint64_t) ((long long) ytop + extra);
  }
  // carry: c = X - (A - B) mod m, 0 or 1
  State probe = st;
  probe.c = 0;
  const Big X0 = state_to_big(probe);
  bool same = true;
  for (int i = 0; i < kLimbs; ++i) same = same && (X0.l[i] == X.l[i]);
  st.c = same ? 0u : 1u;
  return st;
}

// One round of 12 word steps in place (x stays in time order with x[0] oldest afterwards).
GX_HD void swb_round(State &st) {
  uint64_t c = st.c;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const uint64_t y = st.x[(i + 7) % 12] - st.x[i] - c;
    c = y >> 63;
    st.x[i] = y & kMask48;
  }
  st.c = (uint32_t) c;
}

#ifdef __CUDA_ARCH__
// The same round with the 48-bit borrow kept in the hardware carry flag. A word is held as (hi, lo) = (x >> 16,
// (x & 0xffff) << 16): subtracting the two halves with sub.cc/subc.cc makes the borrow out of the upper half exactly the
// borrow of the 48-bit subtraction, and the incoming borrow enters at bit 0 of `lo`, ripples through its 16 zero bits
// (leaving 0xffff there when it was set) and lands on bit 16 -- the word's own bit 0. One `and` clears those 16 bits.
// 3 instructions per word step instead of 11 (two 64-bit subtractions, shift, mask, carry extraction); the generator
// advance is ~70 % of all instructions the noise kernels issue.
__device__ __forceinline__ void swb_round_cc(uint32_t (&lo)[12], uint32_t (&hi)[12], uint32_t &c) {
  asm volatile(
      "{\n\t.reg .u32 t;\n\t"
      "sub.cc.u32 t, 0, %24;\n\t"  // carry flag <- (c != 0)
      "subc.cc.u32 %0, %7, %0;\n\t subc.cc.u32 %12, %19, %12;\n\t and.b32 %0, %0, 0xFFFF0000;\n\t"
      "subc.cc.u32 %1, %8, %1;\n\t subc.cc.u32 %13, %20, %13;\n\t and.b32 %1, %1, 0xFFFF0000;\n\t"
      "subc.cc.u32 %2, %9, %2;\n\t subc.cc.u32 %14, %21, %14;\n\t and.b32 %2, %2, 0xFFFF0000;\n\t"
      "subc.cc.u32 %3, %10, %3;\n\t subc.cc.u32 %15, %22, %15;\n\t and.b32 %3, %3, 0xFFFF0000;\n\t"
      "subc.cc.u32 %4, %11, %4;\n\t subc.cc.u32 %16, %23, %16;\n\t and.b32 %4, %4, 0xFFFF0000;\n\t"
      "subc.cc.u32 %5, %0, %5;\n\t subc.cc.u32 %17, %12, %17;\n\t and.b32 %5, %5, 0xFFFF0000;\n\t"
      "subc.cc.u32 %6, %1, %6;\n\t subc.cc.u32 %18, %13, %18;\n\t and.b32 %6, %6, 0xFFFF0000;\n\t"
      "subc.cc.u32 %7, %2, %7;\n\t subc.cc.u32 %19, %14, %19;\n\t and.b32 %7, %7, 0xFFFF0000;\n\t"
      "subc.cc.u32 %8, %3, %8;\n\t subc.cc.u32 %20, %15, %20;\n\t and.b32 %8, %8, 0xFFFF0000;\n\t"
      "subc.cc.u32 %9, %4, %9;\n\t subc.cc.u32 %21, %16, %21;\n\t and.b32 %9, %9, 0xFFFF0000;\n\t"
      "subc.cc.u32 %10, %5, %10;\n\t subc.cc.u32 %22, %17, %22;\n\t and.b32 %10, %10, 0xFFFF0000;\n\t"
      "subc.cc.u32 %11, %6, %11;\n\t subc.cc.u32 %23, %18, %23;\n\t and.b32 %11, %11, 0xFFFF0000;\n\t"
      "subc.u32 %24, 0, 0;\n\t"  // c <- 0 or 0xffffffff
      "}"
      : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]),
        "+r"(lo[8]), "+r"(lo[9]), "+r"(lo[10]), "+r"(lo[11]), "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]),
        "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7]), "+r"(hi[8]), "+r"(hi[9]), "+r"(hi[10]), "+r"(hi[11]),
        "+r"(c));
}
// The same round with the mask on the FMA pipe: x & 0xFFFF0000 as (x * 2^16 >> 32) * 2^16 = IMAD.HI + IMAD. The alu pipe
// (IADD3.X, LOP3) issues one warp instruction every 2 cycles per scheduler and the generator is nothing but alu work, so
// the noise kernels were bound by that pipe (ncu: alu 71 %, fma idle); this leaves 2 alu + 2 fma instructions per word step.
__device__ __forceinline__ void swb_round_cc_fma(uint32_t (&lo)[12], uint32_t (&hi)[12], uint32_t &c) {
#define GX_SWB_STEP(I, J, HI, HJ) \
  "subc.cc.u32 %" #I ", %" #J ", %" #I ";\n\t subc.cc.u32 %" #HI ", %" #HJ ", %" #HI ";\n\t" \
  "mul.hi.u32 u, %" #I ", 0x10000;\n\t mul.lo.u32 %" #I ", u, 0x10000;\n\t"
  asm volatile(
      "{\n\t.reg .u32 t, u;\n\t"
      "sub.cc.u32 t, 0, %24;\n\t"
      GX_SWB_STEP(0, 7, 12, 19) GX_SWB_STEP(1, 8, 13, 20) GX_SWB_STEP(2, 9, 14, 21) GX_SWB_STEP(3, 10, 15, 22)
      GX_SWB_STEP(4, 11, 16, 23) GX_SWB_STEP(5, 0, 17, 12) GX_SWB_STEP(6, 1, 18, 13) GX_SWB_STEP(7, 2, 19, 14)
      GX_SWB_STEP(8, 3, 20, 15) GX_SWB_STEP(9, 4, 21, 16) GX_SWB_STEP(10, 5, 22, 17) GX_SWB_STEP(11, 6, 23, 18)
      "subc.u32 %24, 0, 0;\n\t"
      "}"
      : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]),
        "+r"(lo[8]), "+r"(lo[9]), "+r"(lo[10]), "+r"(lo[11]), "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]),
        "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7]), "+r"(hi[8]), "+r"(hi[9]), "+r"(hi[10]), "+r"(hi[11]),
        "+r"(c));
#undef GX_SWB_STEP
}

// Generator state kept in the (lo, hi) form across increments (no 64-bit repacking per block).
struct StateCC {
  uint32_t lo[12], hi[12], c;
  __device__ __forceinline__ void load(const uint64_t (&x)[12], uint32_t carry) {
#pragma unroll
    for (int i = 0; i < 12; ++i) { lo[i] = (uint32_t) (x[i] << 16); hi[i] = (uint32_t) (x[i] >> 16); }
    c = carry;
  }
  // increment_state: 33 rounds + 1 word step, then rotate so that word 0 is the oldest again
  template <bool FMA> __device__ __forceinline__ void advance() {
#pragma unroll 3
    for (int r = 0; r < 33; ++r) { if (FMA) swb_round_cc_fma(lo, hi, c); else swb_round_cc(lo, hi, c); }
    uint32_t l0 = lo[0], h0 = hi[0];
    asm volatile("{\n\t.reg .u32 t;\n\t sub.cc.u32 t, 0, %2;\n\t subc.cc.u32 %0, %3, %0;\n\t subc.cc.u32 %1, %4, %1;\n\t"
                 "and.b32 %0, %0, 0xFFFF0000;\n\t subc.u32 %2, 0, 0;\n\t}"
                 : "+r"(l0), "+r"(h0), "+r"(c) : "r"(lo[7]), "r"(hi[7]));
#pragma unroll
    for (int i = 0; i < 11; ++i) { lo[i] = lo[i + 1]; hi[i] = hi[i + 1]; }
    lo[11] = l0; hi[11] = h0;
  }
  // out[2i] = low 24 bits of word i, out[2i+1] = high 24 bits
  __device__ __forceinline__ uint32_t out(int j) const {
    const int i = j >> 1;
    return (j & 1) ? (hi[i] >> 8) : (((hi[i] & 0xFFu) << 16) | (lo[i] >> 16));
  }
};
// The same generator on the FP64 pipe. B200's integer (alu) pipe and its FP64 pipe are equally wide (64 lanes per SM; DFMA
// peak measured 63.7 per clock per SM, profiles/r2s19_fp64_peak.txt) and the noise generation is bound by the alu pipe alone
// (ncu: alu 78 %, fp64 3 %). GSL itself runs ranlxs in doubles; a word is an integer below 2^48 held exactly in a double:
//     y = (x_j - x_i) - c;   c' = (y < 0) ? 1.0 : 0.0;   x_i = c' * 2^48 + y
// = 3 DADD/DFMA + 1 DSETP on the FP64 pipe and ONE alu instruction (the select that forms the high word of c') per word
// step, against 3 alu instructions in the carry-flag form. A fraction of the kernel's CTAs runs this form so that both pipes
// work (rng.cu GenParams::f64_share).
struct StateF64 {
  double x[12], c;
  __device__ __forceinline__ void load(const uint64_t (&w)[12], uint32_t carry) {
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = (double) w[i];
    c = carry ? 1.0 : 0.0;
  }
  __device__ __forceinline__ static void step(double &xi, double xj, double &c) {
    const double y = (xj - xi) - c;
    c = __hiloint2double((y < 0.0) ? 0x3FF00000 : 0, 0);
    xi = fma(c, 281474976710656.0, y);
  }
  __device__ __forceinline__ void advance() {
#pragma unroll 3
    for (int r = 0; r < 33; ++r) {
#pragma unroll
      for (int i = 0; i < 12; ++i) step(x[i], x[(i + 7) % 12], c);
    }
    double x0 = x[0];
    step(x0, x[7], c);
#pragma unroll
    for (int i = 0; i < 11; ++i) x[i] = x[i + 1];
    x[11] = x0;
  }
  __device__ __forceinline__ uint32_t out(int j) const {
    const unsigned long long u = (unsigned long long) x[j >> 1];
    return (j & 1) ? (uint32_t) (u >> 24) : (uint32_t) (u & 0xFFFFFFull);
  }
};
#endif

// increment_state of ranlxs2: 397 = 33*12 + 1 word steps, then rotate so x[0] is the oldest again.
GX_HD void advance_block(State &st) {
#ifdef __CUDA_ARCH__
  uint32_t lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { lo[i] = (uint32_t) (st.x[i] << 16); hi[i] = (uint32_t) (st.x[i] >> 16); }
  uint32_t c = st.c;
#pragma unroll 3
  for (int r = 0; r < 33; ++r) swb_round_cc(lo, hi, c);
#pragma unroll
  for (int i = 0; i < 12; ++i) st.x[i] = ((uint64_t) hi[i] << 16) | (uint64_t) (lo[i] >> 16);
  st.c = c & 1u;
#else
  for (int r = 0; r < 33; ++r) swb_round(st);
#endif
  const uint64_t y = st.x[7] - st.x[0] - (uint64_t) st.c;
  st.c = (uint32_t) (y >> 63);
  const uint64_t newest = y & kMask48;
#pragma unroll
  for (int i = 0; i < 11; ++i) st.x[i] = st.x[i + 1];
  st.x[11] = newest;
}

// the 24 outputs released after an increment: out[2i] = low 24 bits of word i, out[2i+1] = high 24 bits
GX_HD uint32_t block_output(const State &st, int j) {
  const uint64_t w = st.x[j >> 1];
  return (j & 1) ? (uint32_t) (w >> 24) : (uint32_t) (w & kMask24);
}

// ---- ziggurat attempt ---------------------------------------------------------------------------------
struct Attempt {
  int consumed;     // 2, 3 or 4 raw outputs
  bool accepted;
  bool tail;        // accepted value came from the tail branch (x depends on log(); host may re-derive it)
  double value;     // sign * x
  uint32_t tail_u1; // raw integer behind U1 for tail attempts
};

#ifdef __CUDA_ARCH__
#define GX_DMUL(a, b) __dmul_rn(a, b)
#define GX_DADD(a, b) __dadd_rn(a, b)
#else
#define GX_DMUL(a, b) ((a) * (b))
#define GX_DADD(a, b) ((a) + (b))
#endif

// r0..r3: the next four raw outputs (only the first `consumed` are used)
GX_HD Attempt zig_attempt(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, const unsigned int *ktab,
                          const double *wtab, const double *ytab) {
  Attempt a;
  a.tail = false;
  a.tail_u1 = 0;
  uint32_t i = r0 & 0xFFu;
  const uint32_t j = r1 & kMask24;
  const double sign = (i & 0x80u) ? 1.0 : -1.0;
  i &= 0x7fu;
  double x = GX_DMUL((double) j, wtab[i]);
  if (j < ktab[i]) {
    a.consumed = 2; a.accepted = true; a.value = GX_DMUL(GX_DMUL(sign, 1.0), x);
    return a;
  }
  double y;
  const double inv24 = 1.0 / 16777216.0;
  if (i < 127) {
    const double y0 = ytab[i], y1 = ytab[i + 1];
    const double U1 = (double) r2 * inv24;
    y = GX_DADD(y1, GX_DMUL(GX_DADD(y0, -y1), U1));
    a.consumed = 3;
  } else {
    const double R = 3.44428647676;
    const double U1 = 1.0 - (double) r2 * inv24;
    const double U2 = (double) r3 * inv24;
    x = GX_DADD(R, -(log(U1) / R));
    y = GX_DMUL(exp(GX_DMUL(-R, GX_DADD(x, -GX_DMUL(0.5, R)))), U2);
    a.consumed = 4;
    a.tail = true;
    a.tail_u1 = r2;
  }
  a.accepted = y < exp(GX_DMUL(GX_DMUL(-0.5, x), x));
  a.value = GX_DMUL(GX_DMUL(sign, 1.0), x);
  return a;
}

}  // namespace rng
}  // namespace gx

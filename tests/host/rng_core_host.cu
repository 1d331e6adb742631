// Host build of genetic_b200/csrc/rng_core.cuh for the CPU test-suite (tests/test_rng_host.py):
// exercises the skip-ahead algebra and the ziggurat parse without a GPU.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../genetic_b200/csrc/rng_core.cuh"
#include "../../oracle/zig_tables.h"

using namespace gx::rng;

static State state_at_block(unsigned long seed, unsigned long long t) {
  // state after t increments, by skip-ahead
  State s0 = seed_state(seed);
  if (t == 0) return s0;
  Big X0 = state_to_big(s0);
  Big G = modpow(big_a(), 397ull * t);
  return big_to_state(modmul(X0, G));
}

extern "C" {

// 0 on success; checks the LCG equivalence, the state round trip and the jump against plain stepping
int host_rng_selfcheck(unsigned long seed) {
  State st = seed_state(seed);
  // a word step == multiplication by a
  Big a = big_a();
  for (int it = 0; it < 50; ++it) {
    Big X = state_to_big(st);
    State rt = big_to_state(X);
    // the round trip is exact once the state is on the recurrent orbit (after >= 12 steps the carry is consistent)
    if (it > 2) {
      for (int i = 0; i < 12; ++i) if (rt.x[i] != st.x[i]) return 100 + it;
      if (rt.c != st.c) return 200 + it;
    }
    // one word step, time-ordered
    State nx;
    uint64_t y = st.x[7] - st.x[0] - (uint64_t) st.c;
    nx.c = (uint32_t) (y >> 63);
    for (int i = 0; i < 11; ++i) nx.x[i] = st.x[i + 1];
    nx.x[11] = y & kMask48;
    Big Xn = state_to_big(nx);
    Big Xp = modmul(a, X);
    if (memcmp(Xn.l, Xp.l, sizeof(Xn.l)) != 0) return 300 + it;
    st = nx;
  }
  // jump vs stepping
  State s = seed_state(seed);
  for (int t = 1; t <= 40; ++t) {
    advance_block(s);
    if (t == 1 || t == 7 || t == 12 || t == 13 || t == 40) {
      State j = state_at_block(seed, t);
      for (int i = 0; i < 12; ++i) if (j.x[i] != s.x[i]) return 400 + t;
      if (j.c != s.c) return 500 + t;
    }
  }
  return 0;
}

// n raw outputs starting at raw offset `offset` of the stream seeded with `seed` (uses skip-ahead)
void host_rng_raw(unsigned long seed, unsigned long long offset, long n, uint32_t *out) {
  unsigned long long blk = offset / 24;  // completed increments before the block containing `offset`
  State st = state_at_block(seed, blk);
  int j = (int) (offset % 24);
  advance_block(st);
  for (long i = 0; i < n; ++i) {
    out[i] = block_output(st, j);
    if (++j == 24) { j = 0; advance_block(st); }
  }
}

// gaussians by parsing a raw buffer with zig_attempt; returns raws consumed
long host_rng_gaussians_from_raw(const uint32_t *raw, long nraw, long ngauss, double *out) {
  long pos = 0, g = 0;
  while (g < ngauss && pos + 4 <= nraw) {
    Attempt a = zig_attempt(raw[pos], raw[pos + 1], raw[pos + 2], raw[pos + 3], gx_zig_ktab, gx_zig_wtab, gx_zig_ytab);
    pos += a.consumed;
    if (a.accepted) out[g++] = a.value;
  }
  return g == ngauss ? pos : -1;
}
}

// Host build of genetic_b200/csrc/rng_staged_core.cuh for the CPU test-suite (tests/test_rng_host.py): the staged noise
// path (gen -> scan -> resolve -> call-driven delivery -> tail patch) run thread by thread on the CPU, exactly the code
// the kernels of rng.cu execute, against the oracle's sequential GSL restatement. No GPU needed.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../genetic_b200/csrc/rng_staged_core.cuh"
#include "../../oracle/zig_tables.h"

using namespace gx::rng;

struct HostRing { uint32_t *r; uint32_t &at(int s) const { return r[s]; } };
struct HostSink {
  std::vector<StagedTail> *v; int chunk;
  void push(int list, int idx, unsigned int u1, bool positive) {
    StagedTail t; t.chunk = chunk; t.u1 = u1; t.idx = (unsigned short) idx; t.list = (unsigned char) list; t.positive = positive;
    v->push_back(t);
  }
};

static long long raw_budget_host(long long n_gauss) {
  return (long long) (2.058 * (double) n_gauss + 12.0 * sqrt((double) n_gauss) + 128.0);
}

extern "C" {

// Parallel Fourier order (one stream per shell, seed + ks) into out[n][n][n/2+1] complex doubles; `written` (bytes, same
// shape) counts the stores per slot. chunk_len_override < kChunkLen is not supported (the chunk length is a constant).
// Returns 0, or a negative code on an internal inconsistency.
int host_staged_noise(int n, unsigned long seed, int reverse, double *out, unsigned char *written, long long *stats) {
  const int nc = n / 2 + 1;
  const double sigma = 1.0 / sqrt(2.0);
  const Big a = big_a();
  std::vector<StreamDesc> sd;
  int n_chunks = 0;
  for (int ks = 1; ks < n / 2; ++ks) {
    StreamDesc d;
    d.s0 = seed_state(seed + (unsigned long) ks);
    d.b0 = 0; d.entry0 = 0; d.n_gauss = 24ll * ks * ks; d.shell = ks;
    d.xs = state_to_big(d.s0);
    d.first_chunk = n_chunks;
    n_chunks += (int) ((raw_budget_host(d.n_gauss) + kChunkLen - 1) / kChunkLen);
    sd.push_back(d);
  }
  (void) a;
  const Big jump = modpow(big_a(), 397ull * kBc);
  std::vector<uint32_t> stage((size_t) n_chunks * kChunkStage, 0xdeadbeefu);
  std::vector<ChunkFn> fn(n_chunks);
  std::vector<ListMeta> meta(n_chunks);
  std::vector<StagedTail> tails;
  long long multi_hop = 0, unmerged = 0;
  // gen
  for (size_t s = 0; s < sd.size(); ++s) {
    const int nch = (s + 1 < sd.size() ? sd[s + 1].first_chunk : n_chunks) - sd[s].first_chunk;
    Big X = sd[s].xs;
    for (int q = 0; q < nch; ++q) {
      const int chunk = sd[s].first_chunk + q;
      State st = (q == 0) ? sd[s].s0 : big_to_state(X);
      X = modmul(X, jump);
      uint32_t ring[kRing], buf[4], inc[kIncCap];
      HostSink sink{&tails, chunk};
      gen_chunk<0>(sd[s], q == 0, st, HostRing{ring}, HostRing{buf}, HostRing{inc}, gx_zig_ktab, gx_zig_wtab, gx_zig_ytab,
                stage.data() + (size_t) chunk * kChunkStage, fn[chunk], meta[chunk], sink);
      for (int e = 1; e < 4; ++e) {
        if (q > 0 && meta[chunk].rep[e] > 0) ++multi_hop;
        if (q > 0 && meta[chunk].rep[e] < 0) ++unmerged;
      }
    }
  }
  // scan
  std::vector<ScanState> scan(n_chunks);
  {
    ScanState s; s.base = 0; s.entry = 0;
    for (int i = 0; i < n_chunks; ++i) { scan[i] = s; s = apply_fn(fn[i], s); }
  }
  // resolve
  std::vector<ChunkLoc> loc(n_chunks);
  for (size_t s = 0; s < sd.size(); ++s) {
    const int nch = (s + 1 < sd.size() ? sd[s + 1].first_chunk : n_chunks) - sd[s].first_chunk;
    for (int q = 0; q < nch; ++q) {
      const int chunk = sd[s].first_chunk + q;
      loc[chunk] = resolve_chunk(meta[chunk], q == 0 ? 0 : scan[chunk].entry, q == 0 ? 0 : scan[chunk].base,
                                 stage.data() + (size_t) chunk * kChunkStage);
    }
  }
  // tails: host log(), spliced into a side table keyed by (stream, g)
  struct Patch { int stream; long long g; double v; };
  std::vector<Patch> patches;
  for (const StagedTail &t : tails) {
    int s = 0;
    while (s + 1 < (int) sd.size() && sd[s + 1].first_chunk <= t.chunk) ++s;
    const bool first = t.chunk == sd[s].first_chunk;
    const int pi = path_index(meta[t.chunk], first ? 0 : scan[t.chunk].entry, t.list, t.idx);
    if (pi < 0) continue;
    const long long g = loc[t.chunk].base + pi;
    if (g >= sd[s].n_gauss) continue;
    const double U1 = 1.0 - (double) t.u1 / 16777216.0;
    const double x = GX_ZIG_R - std::log(U1) / GX_ZIG_R;
    patches.push_back({s, g, t.positive ? x : -x});
  }
  // deliver
  auto draw = [&](int s, long long g) -> double {
    for (const Patch &p : patches) if (p.stream == s && p.g == g) return p.v;
    const int nch = (s + 1 < (int) sd.size() ? sd[s + 1].first_chunk : n_chunks) - sd[s].first_chunk;
    const int c = locate_chunk(loc.data(), sd[s].first_chunk, nch, g, locate_scale(loc.data(), sd[s].first_chunk, nch));
    const int i = (int) (g - loc[c].base);
    return unpack_draw(fetch_draw(loc[c], stage.data() + (size_t) c * kChunkStage, i), gx_zig_wtab);
  };
  auto store = [&](int a1, int b1, int c1, double re, double im) {
    const int ix = a1 < 0 ? a1 + n : a1, iy = b1 < 0 ? b1 + n : b1;
    const size_t slot = (size_t) c1 + (size_t) nc * ((size_t) iy + (size_t) n * ix);
    out[2 * slot] = re; out[2 * slot + 1] = im;
    if (written[slot] < 255) written[slot]++;
  };
  for (int ks = 1; ks < n / 2; ++ks) {
    const int s = ks - 1;
    for (int p = -ks; p < ks; ++p)
      for (int q = -ks; q < ks; ++q) {
        const long long c0 = ((long long) (p + ks) * (2 * ks) + (q + ks)) * 3;
        for (int face = 0; face < 3; ++face) {
          const FaceSlot fs = face_slot(ks, p, q, face);
          if (!fs.write) continue;
          double re, im;
          call_value<double>(draw(s, 2 * (c0 + face)), draw(s, 2 * (c0 + face) + 1), sigma, nullptr, fs.m, reverse, re, im);
          store(fs.a, fs.b, fs.c, re, fs.neg ? -im : im);
          if (fs.second) store(-fs.a, -fs.b, 0, re, fs.neg ? im : -im);
        }
      }
  }
  if (stats) { stats[0] = n_chunks; stats[1] = (long long) tails.size(); stats[2] = (long long) patches.size(); stats[3] = multi_hop; stats[4] = unmerged; }
  // every slot is written exactly once or never; the never-written ones are the analytic set
  for (int ix = 0; ix < n; ++ix)
    for (int iy = 0; iy < n; ++iy)
      for (int iz = 0; iz < nc; ++iz) {
        const size_t slot = (size_t) iz + (size_t) nc * ((size_t) iy + (size_t) n * ix);
        const bool nd = never_drawn(n, ix, iy, iz);
        if (nd && written[slot] != 0) return -10;
        if (!nd && written[slot] != 1) return -11;
      }
  return 0;
}
}

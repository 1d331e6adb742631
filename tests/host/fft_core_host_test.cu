// Host-side emulation of the block FFT in genetic_b200/csrc/fft_core.cuh (no GPU in the build container):
// the per-thread stage functions are executed for every thread t in sequence, stage by stage, which is
// exactly what the __syncthreads-separated device code does. Compared with a long-double naive DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../genetic_b200/csrc/fft_core.cuh"

using namespace gx;
using namespace gx::fft;

template <typename T, int NC, int DIR, int SEL = 0> double run_one() {
  using B = BlockFFT<T, NC, DIR, 0, SEL>;
  using C = cplx<T>;
  std::vector<C> x(NC), out(NC), line(line_pitch<T>(NC) + 8), tw(NC);
  for (int j = 0; j < NC; ++j) {
    long double a = -2.0L * M_PIl * j / NC;
    tw[j].x = (T) cosl(a); tw[j].y = (T) sinl(a);
  }
  srand(NC * 7 + (DIR > 0));
  for (int i = 0; i < NC; ++i) { x[i].x = (T) (rand() / (double) RAND_MAX - 0.5); x[i].y = (T) (rand() / (double) RAND_MAX - 0.5); }
  auto emit = [&](int k, C v) { out[k] = v; };
  // stage 0 for all threads
  for (int t = 0; t < B::P; ++t) {
    C v[B::E];
    for (int i = 0; i < B::E; ++i) v[i] = x[t + i * B::P];
    B::stage0(v, t, line.data(), tw.data(), emit);
  }
  if constexpr (B::S >= 2) for (int t = 0; t < B::P; ++t) B::template stage<1>(t, line.data(), tw.data(), emit);
  if constexpr (B::S >= 3) for (int t = 0; t < B::P; ++t) B::template stage<2>(t, line.data(), tw.data(), emit);
  double maxerr = 0, maxval = 0;
  for (int k = 0; k < NC; ++k) {
    long double sr = 0, si = 0;
    for (int n = 0; n < NC; ++n) {
      long double a = (DIR < 0 ? -2.0L : 2.0L) * M_PIl * ((long long) n * k % NC) / NC;
      long double c = cosl(a), s = sinl(a);
      sr += x[n].x * c - x[n].y * s;
      si += x[n].x * s + x[n].y * c;
    }
    maxerr = fmax(maxerr, fmax(fabs((double) (out[k].x - sr)), fabs((double) (out[k].y - si))));
    maxval = fmax(maxval, fmax(fabs((double) sr), fabs((double) si)));
  }
  return maxerr / maxval;
}

// the natural-order row transform (DitRows): stage A, B and the final decimation-in-time stage, thread by thread
template <typename T, int M, int DIR> double run_dit() {
  using D = DitRows<T, M, DIR>;
  using C = cplx<T>;
  std::vector<C> x(M), out(M), S(D::smem_elems), tw(M);
  for (int j = 0; j < M; ++j) {
    long double a = -2.0L * M_PIl * j / M;
    tw[j].x = (T) cosl(a); tw[j].y = (T) sinl(a);
  }
  srand(M * 13 + (DIR > 0));
  for (int i = 0; i < M; ++i) { x[i].x = (T) (rand() / (double) RAND_MAX - 0.5); x[i].y = (T) (rand() / (double) RAND_MAX - 0.5); }
  std::vector<int> seen(M, 0);
  for (int t = 0; t < D::P; ++t) {
    C v[16];
    for (int e = 0; e < 16; ++e) { v[e] = x[D::input_index(t, e)]; seen[D::input_index(t, e)]++; }
    D::stage_a(v, t, S.data(), tw.data());
  }
  for (int n = 0; n < M; ++n) if (seen[n] != 1) return 1.0;   // every input is owned by exactly one thread
  for (int t = 0; t < D::P; ++t) D::stage_b(t, S.data());
  std::vector<int> emitted(M, 0);
  for (int t = 0; t < D::P; ++t) D::stage_final(t, S.data(), tw.data(), [&](int k, C v) { out[k] = v; emitted[k]++; });
  for (int k = 0; k < M; ++k) if (emitted[k] != 1) return 1.0;
  double maxerr = 0, maxval = 0;
  for (int k = 0; k < M; ++k) {
    long double sr = 0, si = 0;
    for (int n = 0; n < M; ++n) {
      long double a = (DIR < 0 ? -2.0L : 2.0L) * M_PIl * ((long long) n * k % M) / M;
      long double c = cosl(a), s = sinl(a);
      sr += x[n].x * c - x[n].y * s;
      si += x[n].x * s + x[n].y * c;
    }
    maxerr = fmax(maxerr, fmax(fabs((double) (out[k].x - sr)), fabs((double) (out[k].y - si))));
    maxval = fmax(maxval, fmax(fabs((double) sr), fabs((double) si)));
  }
  return maxerr / maxval;
}

template <int M> int check_dit() {
  double e1 = run_dit<double, M, -1>(), e2 = run_dit<double, M, 1>();
  double f1 = run_dit<float, M, -1>(), f2 = run_dit<float, M, 1>();
  int bad = (e1 > 5e-15 || e2 > 5e-15 || f1 > 3e-6 || f2 > 3e-6);
  printf("DitRows M=%4d  f64 fwd %.2e inv %.2e   f32 fwd %.2e inv %.2e %s\n", M, e1, e2, f1, f2, bad ? "FAIL" : "ok");
  return bad;
}

template <int NC> int check() {
  double e1 = run_one<double, NC, -1>(), e2 = run_one<double, NC, 1>();
  double f1 = run_one<float, NC, -1>(), f2 = run_one<float, NC, 1>();
  int bad = (e1 > 5e-15 || e2 > 5e-15 || f1 > 3e-6 || f2 > 3e-6);
  printf("NC=%4d  f64 fwd %.2e inv %.2e   f32 fwd %.2e inv %.2e %s\n", NC, e1, e2, f1, f2, bad ? "FAIL" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  {  // alternative 8-points-per-thread plan of the f64 column passes
    double e1 = run_one<double, 256, -1, 1>(), e2 = run_one<double, 256, 1, 1>();
    int b = (e1 > 5e-15 || e2 > 5e-15);
    printf("NC= 256 (plan 8x8x4) f64 fwd %.2e inv %.2e %s\n", e1, e2, b ? "FAIL" : "ok");
    bad += b;
  }
  bad += check<2>(); bad += check<4>(); bad += check<8>(); bad += check<16>(); bad += check<32>(); bad += check<64>();
  bad += check<128>(); bad += check<256>(); bad += check<512>(); bad += check<1024>(); bad += check<2048>();
  bad += check<4096>();
  bad += check_dit<256>(); bad += check_dit<512>(); bad += check_dit<1024>(); bad += check_dit<2048>();
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}

// Host run of the mixed-radix Stockham stage of fft_generic_core.cuh against a direct O(N^2) DFT in long double.
#include "../../genetic_b200/csrc/fft_generic_core.cuh"
#include <cmath>
#include <cstdio>
#include <vector>
using namespace gx::fftgen;

static double run(int n, int tc, int dir) {
  Plan p;
  if (!make_plan(n, p)) { std::printf("no plan for %d\n", n); return 1; }
  std::vector<C2<double>> tw(n), a((size_t) n * tc), b((size_t) n * tc);
  for (int t = 0; t < n; ++t) { const long double ang = -2.0L * M_PIl * t / n; tw[t] = {(double) cosl(ang), (double) sinl(ang)}; }
  unsigned s = 12345u + n;
  for (auto &v : a) { s = s * 1664525u + 1013904223u; v.x = (double) (s >> 8) / (1 << 24) - 0.5; s = s * 1664525u + 1013904223u; v.y = (double) (s >> 8) / (1 << 24) - 0.5; }
  const std::vector<C2<double>> x = a;
  int Ns = 1;
  C2<double> *in = a.data(), *out = b.data();
  for (int f = 0; f < p.nf; ++f) {
    const int R = p.radix[f];
    for (int j = 0; j < n / R; ++j)
      for (int c = 0; c < tc; ++c) stage_item<double>(in, out, n, tc, R, Ns, tw.data(), dir, j, c);
    Ns *= R;
    std::swap(in, out);
  }
  double worst = 0;
  for (int c = 0; c < tc; ++c)
    for (int k = 0; k < n; ++k) {
      long double re = 0, im = 0;
      for (int t = 0; t < n; ++t) {
        const long double ang = (dir < 0 ? -2.0L : 2.0L) * M_PIl * ((long long) k * t % n) / n;
        const long double cr = cosl(ang), ci = sinl(ang);
        re += x[(size_t) t * tc + c].x * cr - x[(size_t) t * tc + c].y * ci;
        im += x[(size_t) t * tc + c].x * ci + x[(size_t) t * tc + c].y * cr;
      }
      worst = std::fmax(worst, std::fmax(std::fabs((double) re - in[(size_t) k * tc + c].x), std::fabs((double) im - in[(size_t) k * tc + c].y)));
    }
  return worst / std::sqrt((double) n);
}

int main() {
  bool ok = true;
  const int sizes[] = {6, 10, 12, 24, 30, 48, 60, 96, 100, 120, 192, 210, 360, 22, 26, 34, 250};
  for (int n : sizes)
    for (int dir : {-1, 1}) {
      const double e = run(n, 3, dir);
      std::printf("generic n=%d dir=%d err=%.3g\n", n, dir, e);
      ok = ok && e < 2e-14;
    }
  std::printf(ok ? "ALL OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}

// Mixed-radix line transform for grid sizes that are not powers of two (the reference takes any N through FFTW,
// fourier.hpp:521-563; its tests use 8...256, all powers of two, so this is the generality path, not the fast one).
//
// Stockham autosort over the radices of N, TC lines side by side in shared memory ([N][TC] complex, two buffers). One
// stage with radix R and Ns = product of the radices done so far: work item (j, c), j in [0, N/R):
//     k = j mod Ns
//     out[((j div Ns) Ns R + k + p Ns)][c] = sum_q in[j + q N/R][c] W_N^(q (k N/(Ns R) + p N/R))      p = 0 .. R-1
// with W_N^t = exp(-+ 2 pi i t / N) from one table of N entries (host-made in long double). The butterfly is the plain
// O(R^2) sum for every R, so any prime factor works; 4s are taken first to halve the number of passes for the 2-power part.
#pragma once
#include "gx_common.cuh"

namespace gx {
namespace fftgen {

constexpr int kMaxRadices = 24;

struct Plan {
  int n, nf;
  int radix[kMaxRadices];
};

inline bool make_plan(int n, Plan &p) {
  p.n = n; p.nf = 0;
  int m = n;
  while (m % 4 == 0) { p.radix[p.nf++] = 4; m /= 4; }
  for (int r = 2; r <= m; ++r)
    while (m % r == 0) {
      if (p.nf == kMaxRadices) return false;
      p.radix[p.nf++] = r;
      m /= r;
    }
  return m == 1;
}

template <typename T> struct C2 { T x, y; };

// one work item of one stage; `tw` has n entries (cos, -sin), dir = -1 forward (table as is), +1 inverse (conjugate)
template <typename T>
GX_HD void stage_item(const C2<T> *in, C2<T> *out, int n, int tc, int R, int Ns, const C2<T> *tw, int dir, int j, int c) {
  const int k = j % Ns;
  const int stride = n / R;
  const int kstep = k * (n / (Ns * R));
  const int obase = (j / Ns) * Ns * R + k;
  for (int p = 0; p < R; ++p) {
    int step = kstep + p * stride;
    if (step >= n) step -= n;  // kstep < n/R and p * stride < n, so one correction is enough
    int t = 0;
    T ax = 0, ay = 0;
    for (int q = 0; q < R; ++q) {
      const C2<T> v = in[(size_t) (j + q * stride) * tc + c];
      const C2<T> w = tw[t];
      const T wy = dir < 0 ? w.y : -w.y;
      ax += v.x * w.x - v.y * wy;
      ay += v.x * wy + v.y * w.x;
      t += step;
      if (t >= n) t -= n;
    }
    out[(size_t) (obase + p * Ns) * tc + c] = C2<T>{ax, ay};
  }
}

}  // namespace fftgen
}  // namespace gx

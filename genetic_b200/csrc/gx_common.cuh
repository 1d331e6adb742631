// Common helpers for the genetic_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#define GX_HD __host__ __device__ __forceinline__

namespace gx {

// ---- error plumbing: every C-ABI entry returns 0 on success, non-zero on failure and records a message
void set_error(const std::string &msg);
int fail(const std::string &msg);

#define GX_CUDA_CHECK(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      return ::gx::fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " __FILE__ ":" + \
                        std::to_string(__LINE__));                                                   \
    }                                                                                                \
  } while (0)

#define GX_LAUNCH_CHECK()                                                                            \
  do {                                                                                               \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess) {                                                                         \
      return ::gx::fail(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " at " __FILE__ ":" + \
                        std::to_string(__LINE__));                                                   \
    }                                                                                                \
  } while (0)

// ---- scalar/complex type mapping
template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };
template <typename T> using cplx = typename V2<T>::type;

template <typename T> GX_HD cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename C> GX_HD C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> GX_HD C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> GX_HD C cmul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> GX_HD C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }

// is_pow2 / log2
GX_HD bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// Counts kernel launches made by this library (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long) n; }

// Optional per-kernel-class device timing (CUDA events on the launching stream), used by bench.py for the
// roofline line: gx_timing_enable(1); ...; gx_timing_report(). Off by default (no events recorded).
struct TimedScope {
  int idx;
  cudaStream_t st;
  TimedScope(const char *name, cudaStream_t s);
  ~TimedScope();
};
#define GX_TIMED(name, stream) ::gx::TimedScope _gx_timed_scope(name, stream)

// field geometry: padded half-complex layout, N x N x (N/2+1) complex == N x N x (N+2) reals
struct Geo {
  int n;        // grid side
  int nc;       // n/2+1 complex per row
  GX_HD size_t slots() const { return (size_t) n * n * nc; }
  GX_HD size_t reals() const { return 2 * slots(); }
};
inline Geo geo(int n) { Geo g; g.n = n; g.nc = n / 2 + 1; return g; }

// fourier.hpp:110-151 -- storage slot reached by the reference's half-space iteration?
GX_HD bool slot_is_independent(int n, int ix, int iy, int iz) {
  int h = n >> 1;
  if (iz != 0 && iz != h) return true;
  if (ix > h) return false;
  if ((ix == 0 || ix == h) && iy > h) return false;
  return true;
}

GX_HD int signed_mode(int i, int n) { return i <= (n >> 1) ? i : i - n; }

}  // namespace gx

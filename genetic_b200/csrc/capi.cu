// C-ABI: library, memory, host<->device layout copies, FFT entry points.
#include "../../include/genetic_b200.h"
#include "fft.cuh"

namespace gx {
const char *last_error_cstr();
template <typename T> int mirror(T *field, int n, cudaStream_t st);
int pool_malloc(void **ptr, size_t bytes);  // mempool.cu
int pool_free(void *ptr);
namespace fft { template <typename T> int fft3d_any(T *field, int n, int dir, cudaStream_t st); }  // fft_generic.cu
}

using namespace gx;

// even sizes: powers of two take the tuned passes, the rest the mixed-radix passes (fft_generic.cu)
static int check_n(int n) {
  if (n < 4 || n > 4096 || (n & 1)) return gx::fail("grid size must be even and in [4,4096]");
  return 0;
}
static size_t esize(int dtype) { return dtype == GX_F64 ? 8 : 4; }

extern "C" {

const char *gx_last_error(void) { return gx::last_error_cstr(); }
int gx_version(void) { return 100; }

int gx_device_count(int *count) {
  GX_CUDA_CHECK(cudaGetDeviceCount(count));
  return 0;
}

// The library keeps grow-only workspaces (noise staging lists, |k|^2 table, partial sums, cross-grid tables, twiddles) on the
// device that was current when they were created, and is driven by ONE host thread per process (as the reference is:
// every call comes from its main thread). Binding another device drops them; they are rebuilt there on demand.
int gx_set_device(int device) {
  int cur = -1;
  const bool had = cudaGetDevice(&cur) == cudaSuccess;
  static int bound = -1;
  if (bound >= 0 && bound != device) {
    if (had && cur != bound) GX_CUDA_CHECK(cudaSetDevice(bound));
    gx_workspace_release();
  }
  GX_CUDA_CHECK(cudaSetDevice(device));
  bound = device;
  return gx::fft::ensure_twiddles();
}

int gx_synchronize(void *stream) {
  GX_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t) stream));
  return 0;
}

unsigned long long gx_launch_count(void) { return gx::g_launch_count; }

size_t gx_field_bytes(int n, int dtype) { return (size_t) 2 * n * n * (n / 2 + 1) * esize(dtype); }

int gx_malloc(void **ptr, size_t bytes) { return gx::pool_malloc(ptr, bytes); }
int gx_free(void *ptr) { return gx::pool_free(ptr); }
int gx_memset_zero(void *ptr, size_t bytes, void *stream) {
  GX_CUDA_CHECK(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t) stream));
  return 0;
}
int gx_host_alloc_pinned(void **ptr, size_t bytes) {
  GX_CUDA_CHECK(cudaMallocHost(ptr, bytes));
  return 0;
}
int gx_host_free_pinned(void *ptr) {
  GX_CUDA_CHECK(cudaFreeHost(ptr));
  return 0;
}
int gx_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  GX_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
  return 0;
}
int gx_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  GX_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
  return 0;
}
int gx_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  GX_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  return 0;
}

int gx_field_upload(void *dev_field, const void *host, int n, int dtype, int layout, void *stream) {
  if (check_n(n)) return 1;
  const size_t es = esize(dtype);
  if (layout == GX_LAYOUT_FOURIER) {
    GX_CUDA_CHECK(cudaMemcpyAsync(dev_field, host, gx_field_bytes(n, dtype), cudaMemcpyHostToDevice, (cudaStream_t) stream));
  } else if (layout == GX_LAYOUT_REAL) {
    // rows of n reals on the host -> rows of n+2 reals on the device
    GX_CUDA_CHECK(cudaMemcpy2DAsync(dev_field, (size_t) (n + 2) * es, host, (size_t) n * es, (size_t) n * es,
                                    (size_t) n * n, cudaMemcpyHostToDevice, (cudaStream_t) stream));
  } else {
    return gx::fail("bad layout");
  }
  return 0;
}

int gx_field_download(void *host, const void *dev_field, int n, int dtype, int layout, void *stream) {
  if (check_n(n)) return 1;
  const size_t es = esize(dtype);
  if (layout == GX_LAYOUT_FOURIER) {
    GX_CUDA_CHECK(cudaMemcpyAsync(host, dev_field, gx_field_bytes(n, dtype), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  } else if (layout == GX_LAYOUT_REAL) {
    GX_CUDA_CHECK(cudaMemcpy2DAsync(host, (size_t) n * es, dev_field, (size_t) (n + 2) * es, (size_t) n * es,
                                    (size_t) n * n, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  } else {
    return gx::fail("bad layout");
  }
  return 0;
}

int gx_fft_forward(void *field, int n, int dtype, void *stream) {
  if (check_n(n)) return 1;
  if (dtype == GX_F64) return gx::fft::fft3d_any<double>((double *) field, n, -1, (cudaStream_t) stream);
  if (dtype == GX_F32) return gx::fft::fft3d_any<float>((float *) field, n, -1, (cudaStream_t) stream);
  return gx::fail("bad dtype");
}

int gx_fft_inverse(void *field, int n, int dtype, void *stream) {
  if (check_n(n)) return 1;
  int rc;
  if (dtype == GX_F64) {
    rc = gx::mirror<double>((double *) field, n, (cudaStream_t) stream);
    return rc ? rc : gx::fft::fft3d_any<double>((double *) field, n, 1, (cudaStream_t) stream);
  }
  if (dtype == GX_F32) {
    rc = gx::mirror<float>((float *) field, n, (cudaStream_t) stream);
    return rc ? rc : gx::fft::fft3d_any<float>((float *) field, n, 1, (cudaStream_t) stream);
  }
  return gx::fail("bad dtype");
}

}  // extern "C"

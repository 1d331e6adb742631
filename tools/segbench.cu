// segbench -- what HBM bandwidth does a strided-tile pass reach as a function of the contiguous segment width?
//
// Design probe for the column passes of the 3-D FFT (DESIGN.md "FFT passes"): a CTA owns a tile of R rows x W
// bytes (row stride S bytes), reads it, exchanges it through shared memory (the stand-in for the butterflies)
// and writes it back in place. Variants: LDG -> registers -> STS (what the first kernels did) and
// cp.async.bulk global->shared / shared->global with an mbarrier (TMA bulk engine, UBLKCP in SASS).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/segbench tools/segbench.cu
//   tools/segbench            (prints one line per configuration)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// tile t: column-tile index ct = t % ntile_cols, row-block rb = t / ntile_cols (rb > 0 only when R < total rows:
// emulates one CTA of a cluster that splits the line), batch handled by the caller through base offsets.
struct Params {
  char *base;
  long long row_stride;    // bytes between consecutive rows of a line
  long long ncols_bytes;   // bytes of contiguous columns available per row (tiles = ncols_bytes / W)
  int rows_total;          // line length (rows)
  int R;                   // rows per CTA
  int W;                   // bytes per row segment
  long long batch_stride;  // bytes between batches
  int nbatch;
};

template <int EPT>
__global__ void __launch_bounds__(512) ldg_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2 *s = reinterpret_cast<double2 *>(smem);
  const int wcols = p.W / 16;
  const int nthreads = blockDim.x;
  const long long tiles_per_row = p.ncols_bytes / p.W;
  const int rblocks = p.rows_total / p.R;
  long long t = blockIdx.x;
  const long long ct = t % tiles_per_row; t /= tiles_per_row;
  const int rb = (int) (t % rblocks); t /= rblocks;
  const long long batch = t;
  char *tile = p.base + batch * p.batch_stride + (long long) rb * p.R * p.row_stride + ct * p.W;
  double2 v[EPT];
  const int c = threadIdx.x % wcols, r0 = threadIdx.x / wcols, rstep = nthreads / wcols;
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int r = r0 + i * rstep;
    v[i] = *reinterpret_cast<const double2 *>(tile + (long long) r * p.row_stride + c * 16);
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) s[(r0 + i * rstep) * wcols + c] = v[i];
  __syncthreads();
  const int total = p.R * wcols;
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    double2 a = s[(threadIdx.x + i * nthreads + 17) % total];
    v[i].x += a.y * 1e-300; v[i].y += a.x * 1e-300;
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int r = r0 + i * rstep;
    *reinterpret_cast<double2 *>(tile + (long long) r * p.row_stride + c * 16) = v[i];
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) bulk_kernel(Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const long long tiles_per_row = p.ncols_bytes / p.W;
  const int rblocks = p.rows_total / p.R;
  long long t = blockIdx.x;
  const long long ct = t % tiles_per_row; t /= tiles_per_row;
  const int rb = (int) (t % rblocks); t /= rblocks;
  const long long batch = t;
  char *tile = p.base + batch * p.batch_stride + (long long) rb * p.R * p.row_stride + ct * p.W;
  const uint32_t mb = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(p.R * p.W) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int r = threadIdx.x; r < p.R; r += 32) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + (size_t) r * p.W)),
                   "l"(tile + (long long) r * p.row_stride), "r"(p.W), "r"(mb)
                   : "memory");
    }
  }
  // wait phase 0
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mb), "r"(0) : "memory");
    }
  }
  // touch the tile (stand-in for the butterflies): every thread rewrites a few elements
  double2 *s = reinterpret_cast<double2 *>(smem);
  const int total = p.R * p.W / 16;
  for (int i = threadIdx.x; i < total; i += blockDim.x) { double2 a = s[i]; a.x += a.y * 1e-300; s[i] = a; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int r = threadIdx.x; r < p.R; r += 32) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tile + (long long) r * p.row_stride),
                   "r"(smem_u32(smem + (size_t) r * p.W)), "r"(p.W)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

struct Cfg { const char *name; int R, W, threads, ept; int bulk; };

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2048;
  const long long nc = n / 2 + 1;
  const long long fbytes = (long long) n * n * nc * 16;
  // limit the footprint: x-pattern needs the whole field (stride n*nc*16); use it for both patterns
  char *buf;
  CK(cudaMalloc(&buf, fbytes));
  CK(cudaMemset(buf, 0, fbytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<Cfg> cfgs = {
      {"ldg R=n   W=32", n, 32, 256, 16, 0},     {"ldg R=n   W=64", n, 64, 512, 16, 0},
      {"ldg R=n/2 W=128", n / 2, 128, 512, 16, 0}, {"ldg R=n/4 W=128", n / 4, 128, 256, 16, 0},
      {"ldg R=n/4 W=256", n / 4, 256, 512, 16, 0}, {"ldg R=n/8 W=256", n / 8, 256, 256, 16, 0},
      {"ldg R=n/8 W=512", n / 8, 512, 512, 16, 0}, {"ldg R=n/8 W=256 e8", n / 8, 256, 512, 8, 0},
      {"ldg R=n/8 W=128 e8", n / 8, 128, 256, 8, 0},
      {"bulk R=n   W=64", n, 64, 256, 0, 1},     {"bulk R=n/2 W=128", n / 2, 128, 256, 0, 1},
      {"bulk R=n/4 W=128", n / 4, 128, 256, 0, 1}, {"bulk R=n/4 W=256", n / 4, 256, 256, 0, 1},
      {"bulk R=n/8 W=256", n / 8, 256, 256, 0, 1}, {"bulk R=n/8 W=512", n / 8, 512, 256, 0, 1},
      {"bulk R=n/8 W=128", n / 8, 128, 256, 0, 1}, {"bulk R=n/16 W=256", n / 16, 256, 256, 0, 1},
      {"bulk R=n/16 W=512", n / 16, 512, 256, 0, 1}, {"bulk R=n/16 W=1024", n / 16, 1024, 256, 0, 1},
  };
  printf("n=%d field %.2f GB; traffic per run = 2 x field (subsampled where noted)\n", n, fbytes / 1e9);
  for (int pattern = 0; pattern < 2; ++pattern) {
    for (auto &c : cfgs) {
      Params p;
      p.base = buf; p.rows_total = n; p.R = c.R; p.W = c.W;
      long long usable;  // contiguous bytes per row that are whole tiles
      if (pattern == 0) {  // x-pass: row stride n*nc*16, columns = n*nc*16 bytes contiguous
        p.row_stride = n * nc * 16; p.ncols_bytes = (n * nc * 16 / c.W) * c.W; p.batch_stride = 0; p.nbatch = 1;
      } else {  // y-pass: row stride nc*16, columns nc*16 bytes, batch = x planes
        p.row_stride = nc * 16; p.ncols_bytes = (nc * 16 / c.W) * c.W; p.batch_stride = n * nc * 16; p.nbatch = n;
      }
      usable = p.ncols_bytes;
      const long long tiles = (usable / c.W) * (n / c.R) * p.nbatch;
      const size_t smem = (size_t) c.R * c.W;
      if (c.bulk) CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      else {
        if (c.ept == 16) CK(cudaFuncSetAttribute(ldg_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        else CK(cudaFuncSetAttribute(ldg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        if ((long long) c.threads * c.ept != (long long) c.R * c.W / 16) { printf("%-22s bad cfg\n", c.name); continue; }
      }
      float best = 1e30f;
      for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        if (c.bulk) bulk_kernel<<<(unsigned) tiles, c.threads, smem>>>(p);
        else if (c.ept == 16) ldg_kernel<16><<<(unsigned) tiles, c.threads, smem>>>(p);
        else ldg_kernel<8><<<(unsigned) tiles, c.threads, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      const double bytes = 2.0 * (double) tiles * c.R * c.W;
      printf("%s %-22s smem %6zu KB tiles %9lld  %8.3f ms  %7.1f GB/s\n", pattern == 0 ? "X" : "Y", c.name, smem / 1024,
             tiles, best, bytes / best / 1e6);
      fflush(stdout);
    }
  }
  return 0;
}

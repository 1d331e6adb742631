// FP64 pipe peak of this GPU: independent DFMA chains, 8 per thread, enough warps to cover the latency.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cuda_runtime.h>
#include <cstdio>
__global__ void dfma(double *out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 1 << 16;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = (double) blocks * threads * 8.0 * iters;
  printf("fp64_peak: %d SMs, %.3f ms, %.2f TFLOP/s (DFMA = 2 flop), %.1f DFMA/clk/SM at 1.965 GHz\n", sms, best,
         2.0 * fmas / (best * 1e-3) / 1e12, fmas / (best * 1e-3) / sms / 1.965e9);
  return 0;
}

// Micro-benchmark: sustained FP64 DMMA (mma.sync.m8n8k4.f64) and DFMA throughput of the whole chip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
  const int iters = 20000;
  for (int cps = 1; cps <= 4; cps *= 2) {
    const int grid = sms * cps;
    {
      float ms = time_ms([&] { dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 0.9999999); });
      double flop = (double)grid * 8 /*warps*/ * iters * 8 * 512.0;
      printf("DMMA m8n8k4, 8 accumulators/warp, %d CTA(s)/SM x 8 warps: %.2f TFLOP/s\n", cps, flop / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_kernel<2><<<grid, 256>>>(out, iters * 4, 1.0000001, 0.9999999); });
      double flop = (double)grid * 8 * iters * 4 * 2 * 512.0;
      printf("DMMA m8n8k4, 2 accumulators/warp, %d CTA(s)/SM x 8 warps: %.2f TFLOP/s\n", cps, flop / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_kernel<8><<<grid, 256>>>(out, iters * 4, 1.0000001, 1e-9); });
      double flop = (double)grid * 256 * iters * 4 * 8 * 2.0;
      printf("DFMA, 8 chains/thread, %d CTA(s)/SM x 256 threads: %.2f TFLOP/s\n", cps, flop / ms * 1e-9);
    }
  }
  return 0;
}

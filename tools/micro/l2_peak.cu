// Micro-benchmark: read bandwidth the whole chip sustains out of L2 (working set well below the 126 MB of L2) and the
// rate of scattered 216-byte block gathers (the access pattern of the Schur S product), next to the HBM copy figure of
// MEASURED_PEAKS.json.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_peak l2_peak.cu ; prints GB/s.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// every thread streams 16-byte loads over the buffer, `reps` passes; L1 is bypassed (ld.global.cg)
__global__ void __launch_bounds__(256) l2_read_kernel(const double2* __restrict__ buf, size_t n16, int reps, double* out) {
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      const double2 v = __ldcg(buf + i);
      acc += v.x + v.y;
    }
  if (acc == 12345.678) out[0] = acc;
}

// one warp per "match": 27 lanes read a 216-byte block at a pseudo-random block index (8-byte loads, L1 allowed), like
// the operand gathers of schur_s9_kernel; `span_blocks` sets the working set
__global__ void __launch_bounds__(128) gather_kernel(const double* __restrict__ buf, size_t span_blocks, int per_warp,
                                                     double* out) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned long long s = warp * 0x9E3779B97F4A7C15ull + 12345;
  double acc = 0.0;
  for (int q = 0; q < per_warp; ++q) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    const size_t blk = (s >> 20) % span_blocks;
    if (lane < 27) acc += __ldg(buf + blk * 27 + lane);
  }
  if (acc == 12345.678) out[0] = acc;
}

int main() {
  const size_t bytes = 48ull << 20;  // 48 MB: resident in L2
  double* buf;
  double* out;
  cudaMalloc(&buf, (1ull << 30) + 4096);
  cudaMalloc(&out, 64);
  cudaMemset(buf, 0, (1ull << 30) + 4096);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  {
    const int reps = 40;
    l2_read_kernel<<<148 * 8, 256>>>((const double2*)buf, bytes / 16, 2, out);  // warm L2
    cudaEventRecord(e0);
    l2_read_kernel<<<148 * 8, 256>>>((const double2*)buf, bytes / 16, reps, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"l2_read_gbs\": %.1f, \"working_set_mb\": %zu, ", bytes * (double)reps / ms / 1e6, bytes >> 20);
  }
  {
    // gathers of 216-byte blocks out of 1 GB (the size of the W buffer at Final-shape: mostly L2 misses -> DRAM) and
    // out of 48 MB (L2 hits)
    const int per_warp = 64;
    const int grid = 148 * 64;
    const size_t warps = (size_t)grid * 4;
    for (int pass = 0; pass < 2; ++pass) {
      const size_t span = (pass == 0 ? (1ull << 30) : bytes) / 216;
      gather_kernel<<<grid, 128>>>(buf, span, 8, out);
      cudaEventRecord(e0);
      gather_kernel<<<grid, 128>>>(buf, span, per_warp, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      const double blocks = (double)warps * per_warp;
      printf("\"gather216_%s_gbs\": %.1f, \"gather216_%s_gblocks_per_s\": %.2f%s", pass == 0 ? "dram" : "l2",
             blocks * 216 / ms / 1e6, pass == 0 ? "dram" : "l2", blocks / ms / 1e6, pass == 0 ? ", " : "}\n");
    }
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

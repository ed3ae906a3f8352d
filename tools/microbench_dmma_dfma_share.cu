// Microbenchmark: do DMMA (mma.m8n8k4.f64) and DFMA share one FP64 datapath on this part?  Eight warps per SM; per warp a
// fixed amount of work: either DMMAs on 16 independent accumulator pairs or DFMAs on 16 independent chains.  Three runs:
// the four "DMMA warps" alone (the other four idle), the four "DFMA warps" alone, and both kinds together.  Separate pipes
// would make the combined run take max(t_dmma, t_dfma); a shared datapath makes it their sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_dmma_dfma_share tools/microbench_dmma_dfma_share.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode bit 0: warps 0..3 run DMMAs; bit 1: warps 4..7 run DFMAs
__global__ void __launch_bounds__(256) k(double* out, int mode, int iters) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i][0] = acc[i][1] = 1e-3 * (lane + i);
  const double a = 1.0 + 1e-9 * lane, b = 1e-9;
  if (warp < 4) {
    if (mode & 1)
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(acc[i][0], acc[i][1], a, b);
  } else if (mode & 2) {
    // 256 FMAs per lane-group equal one DMMA's 256 FMAs: 16 DFMA warp instructions (32 lanes each) per DMMA-equivalent x 2
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) {
          acc[i][0] = fma(acc[i][0], a, b);
          acc[i][1] = fma(acc[i][1], a, b);
        }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount, iters = 20000;
  double* out;
  cudaMalloc(&out, (size_t)nsm * 256 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[4] = {0, 0, 0, 0};
  for (int mode = 1; mode <= 3; mode++) {
    k<<<nsm, 256>>>(out, mode, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0); k<<<nsm, 256>>>(out, mode, iters); cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float t; cudaEventElapsedTime(&t, e0, e1);
      if (t < best) best = t;
    }
    ms[mode] = best;
  }
  const double dmma_flops = (double)nsm * 4 * iters * 16 * 512.0, dfma_flops = (double)nsm * 4 * 32 * iters * 8.0 * 16 * 2 * 2;
  printf("{\"test\":\"dmma_dfma_share\",\"dmma_only_ms\":%.3f,\"dmma_only_tflops\":%.2f,\"dfma_only_ms\":%.3f,\"dfma_only_tflops\":%.2f,"
         "\"both_ms\":%.3f,\"sum_ms\":%.3f,\"max_ms\":%.3f}\n",
         ms[1], dmma_flops / ms[1] * 1e-9, ms[2], dfma_flops / ms[2] * 1e-9, ms[3], ms[1] + ms[2], ms[1] > ms[2] ? ms[1] : ms[2]);
  return 0;
}

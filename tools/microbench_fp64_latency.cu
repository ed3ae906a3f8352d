// Dependent-chain latencies of the FP64 building blocks of the diagonal-block factorisation (one warp, clock64):
//   DFMA, rsqrt(double), 1/sqrt via sqrt + div, 64-bit warp shuffle, shared-memory load -> DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_fp64_latency tools/microbench_fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, double x0, int n) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x0 + threadIdx.x * 1e-3;
  __syncthreads();
  double x = x0 + threadIdx.x * 1e-6, y = 1.0000001;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (OP == 0) x = fma(x, y, 1e-9);
      if (OP == 1) x = rsqrt(x) + 1.5;
      if (OP == 2) x = 1.0 / sqrt(x) + 1.5;
      if (OP == 3) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
      if (OP == 4) x = fma(sm[(__double2loint(x) + u) & 63], y, x);
      if (OP == 5) x = __drcp_rn(x) + 1.5;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 32 * 8);
  cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA", "rsqrt(double) + DADD", "1/sqrt(double) + DADD", "shfl 64-bit", "LDS -> DFMA (address dependent)",
                         "__drcp_rn + DADD"};
  const int n = 2000;
  for (int op = 0; op < 6; op++) {
    long long h = 0;
    for (int rep = 0; rep < 2; rep++) {
      switch (op) {
        case 0: chain<0><<<1, 32>>>(out, cyc, 1.0, n); break;
        case 1: chain<1><<<1, 32>>>(out, cyc, 1.0, n); break;
        case 2: chain<2><<<1, 32>>>(out, cyc, 1.0, n); break;
        case 3: chain<3><<<1, 32>>>(out, cyc, 1.0, n); break;
        case 4: chain<4><<<1, 32>>>(out, cyc, 1.0, n); break;
        case 5: chain<5><<<1, 32>>>(out, cyc, 1.0, n); break;
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("{\"op\": \"%s\", \"cycles_per_op\": %.1f}\n", names[op], (double)h / (8.0 * n));
  }
  return 0;
}

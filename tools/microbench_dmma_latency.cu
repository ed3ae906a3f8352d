// Microbenchmark: latency and single-warp issue rate of DMMA (mma.m8n8k4.f64) on one SM -- what bounds the short DMMA
// sections on the critical path of ONE factorisation (36-fragment update of a diagonal block, panel product), where a
// scheduler holds a single warp.  Prints cycles per DMMA for 1..16 independent accumulator chains, operands in
// registers and operands loaded from shared memory right before use (one A and one B fragment per DMMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_dmma_latency tools/microbench_dmma_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH, bool SMEM>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ double sm[64 * 68];
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int e = tid; e < 64 * 68; e += blockDim.x) sm[e] = 1e-3 * (e % 17);
  __syncthreads();
  double acc[CH][2];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c][0] = acc[c][1] = 0.0;
  double a = 1e-3 * lane, b = 2e-3 * lane;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < 16; kk += 4)
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (SMEM) {
          a = sm[((8 * c + g) % 64) * 68 + kk + t + (it & 3) * 16];
          b = sm[((8 * (c + 1) + g) % 64) * 68 + kk + t + (it & 3) * 16];
        }
        dmma884(acc[c][0], acc[c][1], a, b);
      }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CH, bool SMEM>
void run(double* out, long long* cyc, int warps) {
  const int iters = 2000;
  k<CH, SMEM><<<1, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  k<CH, SMEM><<<1, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"test\":\"dmma_latency\",\"chains\":%d,\"smem_operands\":%d,\"warps_on_sm\":%d,\"cycles_per_dmma_per_warp\":%.1f}\n", CH,
         (int)SMEM, warps, (double)h / (iters * 4.0 * CH));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64);
  for (int warps : {1, 4, 8}) {
    run<1, false>(out, cyc, warps);
    run<2, false>(out, cyc, warps);
    run<4, false>(out, cyc, warps);
    run<6, false>(out, cyc, warps);
    run<8, false>(out, cyc, warps);
    run<16, false>(out, cyc, warps);
    run<1, true>(out, cyc, warps);
    run<6, true>(out, cyc, warps);
    run<16, true>(out, cyc, warps);
  }
  return 0;
}

// Microbenchmark: DMMA (mma.m8n8k4.f64) fed from shared memory with the fragment pattern of tile_gemm.cuh
// (warp tile 32x32: 8 LDS.64 per 16 DMMA), at 1..4 CTAs of 4 warps per SM, with and without a CTA barrier per
// 16-deep slab.  Answers: how much of the 37 TF DMMA issue peak is reachable with smem-fed operands and how many
// warps per scheduler it takes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_dmma_smem ...
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MI, int NI, bool BARRIER>
__global__ void __launch_bounds__(128) k_smem(double* out, int iters) {
  extern __shared__ double sm[];
  constexpr int LD = 20, BK = 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  for (int e = tid; e < 2 * 64 * LD * 3; e += 128) sm[e] = 1e-3 * (e % 17);
  __syncthreads();
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < iters; it++) {
    const double* sA = sm + (it % 3) * (2 * 64 * LD);
    const double* sB = sA + 64 * LD;
    if (BARRIER) __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = sA[((wm * 8 * MI + i * 8 + g) % 64) * LD + kk + t];
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = sB[((wn * 8 * NI + j * 8 + g) % 64) * LD + kk + t];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * 128 + tid] = s;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

template <int MI, int NI, bool BARRIER>
void run(int nsm, double* out) {
  const int iters = 20000;
  for (int per_sm = 1; per_sm <= 4; per_sm++) {
    // dynamic smem chosen so that exactly per_sm CTAs fit on one SM
    size_t smem = (size_t)(227 * 1024 / per_sm) - 2048;
    if (smem < 2 * 64 * 20 * 3 * 8) smem = 2 * 64 * 20 * 3 * 8;
    CK(cudaFuncSetAttribute(k_smem<MI, NI, BARRIER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem<MI, NI, BARRIER>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_smem<MI, NI, BARRIER>, 128, smem));
    const int blocks = nsm * per_sm;
    float ms = time_ms([&] { k_smem<MI, NI, BARRIER><<<blocks, 128, smem>>>(out, iters); }, 3);
    double fl = (double)blocks * 4 * iters * 4 * MI * NI * 512.0;
    printf("{\"test\":\"dmma_smem\",\"mi\":%d,\"ni\":%d,\"barrier\":%d,\"ctas_per_sm\":%d,\"occ\":%d,\"tflops\":%.2f}\n", MI, NI,
           (int)BARRIER, per_sm, occ, fl / ms * 1e-9);
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, (size_t)nsm * 8 * 128 * 8));
  run<4, 4, false>(nsm, out);
  run<4, 4, true>(nsm, out);
  run<8, 4, false>(nsm, out);
  run<4, 2, false>(nsm, out);
  return 0;
}

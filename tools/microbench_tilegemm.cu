// Microbenchmark: the TileGemm primitive (tile_gemm.cuh) as a plain DGEMM  C = A B^T  (both K-contiguous),
// M = N = 8192, K = 2048, for several tile / pipeline configurations.  Shows how far the cp.async + DMMA main loop
// is from the 37 TF DMMA issue peak when nothing else (epilogues, flags, tails) is in the way.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../andvaranaut_b200/csrc/tile_gemm.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
using namespace avn;

template <typename G, int MINB>
__global__ void __launch_bounds__(G::NTHREADS, MINB) gemm_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                                 double* __restrict__ C, int M, int N, int K) {
  extern __shared__ double smem[];
  const int tm = blockIdx.x % (M / G::BM), tn = blockIdx.x / (M / G::BM);
  G g;
  g.zero();
  g.run(smem, A + (int64_t)tm * G::BM * K, K, G::BM, B + (int64_t)tn * G::BN * K, K, G::BN, K);
  g.for_each([&](int r, int c, double& v) { C[(int64_t)(tm * G::BM + r) * N + tn * G::BN + c] = v; });
}

__global__ void fill(double* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1e-3 * (i % 1013);
}

template <typename G, int MINB>
void run(const char* name, const double* A, const double* B, double* C, int M, int N, int K) {
  CK(cudaFuncSetAttribute(gemm_kernel<G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_kernel<G, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_kernel<G, MINB>, G::NTHREADS, G::SMEM_BYTES));
  const int grid = (M / G::BM) * (N / G::BN);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gemm_kernel<G, MINB><<<grid, G::NTHREADS, G::SMEM_BYTES>>>(A, B, C, M, N, K);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(a);
    gemm_kernel<G, MINB><<<grid, G::NTHREADS, G::SMEM_BYTES>>>(A, B, C, M, N, K);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  printf("{\"test\":\"tilegemm\",\"cfg\":\"%s\",\"threads\":%d,\"smem_kb\":%.1f,\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", name,
         G::NTHREADS, G::SMEM_BYTES / 1024.0, occ, best, 2.0 * M * (double)N * K / best * 1e-9);
}

int main() {
  const int M = 8192, N = 8192, K = 2048;
  double *A, *B, *C;
  CK(cudaMalloc(&A, (size_t)M * K * 8));
  CK(cudaMalloc(&B, (size_t)N * K * 8));
  CK(cudaMalloc(&C, (size_t)M * N * 8));
  fill<<<1024, 256>>>(A, (size_t)M * K);
  fill<<<1024, 256>>>(B, (size_t)N * K);
  CK(cudaDeviceSynchronize());
  run<TileGemm<64, 64, 16, 32, 32, 3, false, false>, 3>("64x64 bk16 s3 w32x32", A, B, C, M, N, K);
  run<TileGemm<64, 64, 16, 32, 32, 4, false, false>, 2>("64x64 bk16 s4 w32x32", A, B, C, M, N, K);
  run<TileGemm<64, 64, 32, 32, 32, 2, false, false>, 3>("64x64 bk32 s2 w32x32", A, B, C, M, N, K);
  run<TileGemm<64, 64, 16, 32, 32, 2, false, false>, 4>("64x64 bk16 s2 w32x32 (4/SM)", A, B, C, M, N, K);
  run<TileGemm<128, 64, 16, 32, 32, 3, false, false>, 2>("128x64 bk16 s3 w32x32", A, B, C, M, N, K);
  run<TileGemm<128, 64, 16, 32, 32, 4, false, false>, 1>("128x64 bk16 s4 w32x32 (1/SM)", A, B, C, M, N, K);
  run<TileGemm<128, 128, 16, 64, 32, 3, false, false>, 1>("128x128 bk16 s3 w64x32", A, B, C, M, N, K);
  run<TileGemm<128, 128, 16, 32, 32, 3, false, false>, 1>("128x128 bk16 s3 w32x32 (512 thr)", A, B, C, M, N, K);
  run<TileGemm<128, 128, 8, 64, 32, 4, false, false>, 1>("128x128 bk8 s4 w64x32", A, B, C, M, N, K);
  return 0;
}

// Microbenchmark: the k-contiguous operand load of tile_gemm.cuh done by TMA instead of 16-byte cp.async, so that the
// "cp.async, not TMA" choice of the FP64 tile primitive rests on a measurement.
//
// Same DGEMM as microbench_tilegemm.cu (C = A B^T, both operands k-contiguous, M = N = 8192, K = 2048, 64 x 64 CTA tile,
// 4 compute warps with 32 x 32 warp tiles, 16-deep slabs):
//   * operands arrive by cp.async.bulk.tensor.2d (one elected thread, box 16 doubles x 64 rows = 128-byte rows, dense in
//     shared memory with the 128-byte swizzle -- 16 KB per stage instead of 20 KB with the padded pitch), completion on an
//     mbarrier per stage; slots are handed back through a second mbarrier per stage (one arrive per compute warp), so
//     there is NO CTA-wide barrier in the main loop;
//   * the m8n8k4 fragments are read conflict-free from the swizzled tile by giving lane group g the row 2g (g < 4) or
//     2(g-4)+1 of its 8-row fragment: a half-warp then touches rows {0,2,4,6} (or {1,3,5,7}) whose swizzled 16-byte chunks
//     are all different.  The accumulator rows / columns are permuted accordingly (a production kernel would carry that
//     permutation through every epilogue);
//   * variants: the TMA issued by thread 0 of compute warp 0 (128 threads, 3 CTAs per SM) or by a fifth, dedicated
//     producer warp (160 threads; 2 CTAs per SM at 168 registers).
// The result is checked bit for bit against the cp.async kernel (same DMMA order).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/microbench_tilegemm_tma tools/microbench_tilegemm_tma.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../andvaranaut_b200/csrc/tile_gemm.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
using namespace avn;

// ---- reference: the cp.async primitive ----
template <typename G, int MINB>
__global__ void __launch_bounds__(G::NTHREADS, MINB) gemm_cpasync(const double* __restrict__ A, const double* __restrict__ B,
                                                                  double* __restrict__ C, int M, int N, int K) {
  extern __shared__ double smem[];
  const int tm = blockIdx.x % (M / G::BM), tn = blockIdx.x / (M / G::BM);
  G g;
  g.zero();
  g.run(smem, A + (int64_t)tm * G::BM * K, K, G::BM, B + (int64_t)tn * G::BN * K, K, G::BN, K);
  g.for_each([&](int r, int c, double& v) { C[(int64_t)(tm * G::BM + r) * N + tn * G::BN + c] = v; });
}

// ---- TMA variant ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

constexpr int TBK = 16, TSTAGE_BYTES = 2 * 64 * TBK * 8;   // A tile + B tile, dense: 16 KB

__device__ __forceinline__ int rho(int g) { return g < 4 ? 2 * g : 2 * (g - 4) + 1; }
// element (row, k) of a swizzle-128B tile whose rows are 16 doubles: 16-byte chunk index XOR (row mod 8)
__device__ __forceinline__ double lds_swz(const unsigned char* tile, int row, int k) {
  return *reinterpret_cast<const double*>(tile + row * 128 + ((((k >> 1) ^ row) & 7) << 4) + ((k & 1) << 3));
}

template <int STAGES, bool PRODUCER_WARP>
__global__ void __launch_bounds__(PRODUCER_WARP ? 160 : 128, PRODUCER_WARP ? 2 : 3)
    gemm_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, double* __restrict__ C, int M,
             int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tm = blockIdx.x % (M / 64), tn = blockIdx.x / (M / 64);
  const int KT = K / TBK;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int kt) {   // one thread
    const int s = kt % STAGES;
    if (kt >= STAGES) mbar_wait(&empty[s], ((kt / STAGES) - 1) & 1);
    mbar_expect_tx(&full[s], TSTAGE_BYTES);
    tma_load_2d(smem + s * TSTAGE_BYTES, &mapA, kt * TBK, tm * 64, &full[s]);
    tma_load_2d(smem + s * TSTAGE_BYTES + 64 * TBK * 8, &mapB, kt * TBK, tn * 64, &full[s]);
  };
  if (PRODUCER_WARP && warp == 4) {
    if (lane == 0)
      for (int kt = 0; kt < KT; kt++) issue(kt);
    return;
  }
  const int wm = warp % 2, wn = warp / 2, g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  if (!PRODUCER_WARP && tid == 0)
    for (int kt = 0; kt < STAGES - 1 && kt < KT; kt++) issue(kt);
  const int rg = rho(g);
  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % STAGES;
    if (!PRODUCER_WARP && tid == 0 && kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
    mbar_wait(&full[s], (kt / STAGES) & 1);
    const unsigned char* sA = smem + s * TSTAGE_BYTES;
    const unsigned char* sB = sA + 64 * TBK * 8;
#pragma unroll
    for (int kk = 0; kk < TBK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = lds_swz(sA, wm * 32 + i * 8 + rg, kk + t);
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = lds_swz(sB, wn * 32 + j * 8 + rg, kk + t);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = tm * 64 + wm * 32 + i * 8 + rg;
      const int c0 = tn * 64 + wn * 32 + j * 8 + rho(2 * t), c1 = tn * 64 + wn * 32 + j * 8 + rho(2 * t + 1);
      C[(int64_t)r * N + c0] = acc[i][j][0];
      C[(int64_t)r * N + c1] = acc[i][j][1];
    }
}

__global__ void fill(double* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1e-3 * (i % 1013);
}
__global__ void diff_kernel(const double* a, const double* b, size_t n, unsigned long long* ndiff) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (a[i] != b[i]) atomicAdd(ndiff, 1ull);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, double* base, int rows, int K) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 8};
  cuuint32_t box[2] = {TBK, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

template <typename F>
static float best_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

template <int STAGES, bool PW>
static void run_tma(const char* name, const CUtensorMap& mA, const CUtensorMap& mB, double* C, const double* Cref, int M, int N,
                    int K, unsigned long long* ndiff) {
  const size_t smem = (size_t)STAGES * TSTAGE_BYTES + 1024;
  CK(cudaFuncSetAttribute(gemm_tma<STAGES, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(gemm_tma<STAGES, PW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_tma<STAGES, PW>, PW ? 160 : 128, smem));
  const int grid = (M / 64) * (N / 64);
  CK(cudaMemset(C, 0, (size_t)M * N * 8));
  const float ms = best_ms([&] { gemm_tma<STAGES, PW><<<grid, PW ? 160 : 128, smem>>>(mA, mB, C, M, N, K); });
  CK(cudaMemset(ndiff, 0, 8));
  diff_kernel<<<1024, 256>>>(C, Cref, (size_t)M * N, ndiff);
  unsigned long long h = 0;
  CK(cudaMemcpy(&h, ndiff, 8, cudaMemcpyDeviceToHost));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, gemm_tma<STAGES, PW>));
  printf("{\"test\":\"tilegemm_tma\",\"cfg\":\"%s\",\"threads\":%d,\"smem_kb\":%.1f,\"regs\":%d,\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f,"
         "\"elements_differing_from_cp_async\":%llu}\n",
         name, PW ? 160 : 128, smem / 1024.0, fa.numRegs, occ, ms, 2.0 * M * (double)N * K / ms * 1e-9, h);
}

int main() {
  const int M = 8192, N = 8192, K = 2048;
  double *A, *B, *C, *Cref;
  unsigned long long* ndiff;
  CK(cudaMalloc(&A, (size_t)M * K * 8));
  CK(cudaMalloc(&B, (size_t)N * K * 8));
  CK(cudaMalloc(&C, (size_t)M * N * 8));
  CK(cudaMalloc(&Cref, (size_t)M * N * 8));
  CK(cudaMalloc(&ndiff, 8));
  fill<<<1024, 256>>>(A, (size_t)M * K);
  fill<<<1024, 256>>>(B, (size_t)N * K);
  CK(cudaDeviceSynchronize());
  {
    using G = TileGemm<64, 64, 16, 32, 32, 3, false, false>;
    CK(cudaFuncSetAttribute(gemm_cpasync<G, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    CK(cudaFuncSetAttribute(gemm_cpasync<G, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_cpasync<G, 3>, G::NTHREADS, G::SMEM_BYTES));
    const int grid = (M / 64) * (N / 64);
    const float ms = best_ms([&] { gemm_cpasync<G, 3><<<grid, G::NTHREADS, G::SMEM_BYTES>>>(A, B, Cref, M, N, K); });
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, gemm_cpasync<G, 3>));
    printf("{\"test\":\"tilegemm_tma\",\"cfg\":\"cp.async 16 B, padded pitch, 3 stages, __syncthreads per slab (tile_gemm.cuh)\",\"threads\":128,"
           "\"smem_kb\":%.1f,\"regs\":%d,\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",
           G::SMEM_BYTES / 1024.0, fa.numRegs, occ, ms, 2.0 * M * (double)N * K / ms * 1e-9);
  }
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const CUtensorMap mA = make_map(enc, A, M, K), mB = make_map(enc, B, N, K);
  run_tma<3, false>("TMA swizzle-128B + mbarriers, 3 stages, issued by compute warp 0", mA, mB, C, Cref, M, N, K, ndiff);
  run_tma<4, false>("TMA swizzle-128B + mbarriers, 4 stages, issued by compute warp 0", mA, mB, C, Cref, M, N, K, ndiff);
  run_tma<3, true>("TMA swizzle-128B + mbarriers, 3 stages, producer warp", mA, mB, C, Cref, M, N, K, ndiff);
  run_tma<6, true>("TMA swizzle-128B + mbarriers, 6 stages, producer warp", mA, mB, C, Cref, M, N, K, ndiff);
  return 0;
}

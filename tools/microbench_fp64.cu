// Micro-benchmarks that fix the FP64 roofline denominators on the B200 box:
//   DFMA vector pipe, DMMA (mma.sync f64) shapes, cuBLAS DGEMM, cuSOLVER potrf (the library bar).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = fma(r[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma884_kernel(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma1688_kernel(double* out, int iters) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 - threadIdx.x * 1e-4, b1 = b0 * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma16816_kernel(double* out, int iters) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1.0 - threadIdx.x * 1e-4 * (i + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void fill_kernel(double* p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v + i;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", prop.name, nsm, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
  const int iters = 4096;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32, blocks = nsm * 2;
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("{\"test\":\"dfma\",\"warps_per_cta\":%d,\"ctas\":%d,\"tflops\":%.3f}\n", warps, blocks, fl / ms * 1e-9);
  }
  for (int warps = 4; warps <= 16; warps *= 2) {
    int threads = warps * 32, blocks = nsm * 2;
    float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
    double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks;
    printf("{\"test\":\"dmma_m8n8k4\",\"warps_per_cta\":%d,\"tflops\":%.3f}\n", warps, fl / ms * 1e-9);
    ms = time_ms([&] { dmma1688_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
    fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)warps * blocks;
    printf("{\"test\":\"dmma_m16n8k8\",\"warps_per_cta\":%d,\"tflops\":%.3f}\n", warps, fl / ms * 1e-9);
    ms = time_ms([&] { dmma16816_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
    fl = 2.0 * 16 * 8 * 16 * 8 * iters * (double)warps * blocks;
    printf("{\"test\":\"dmma_m16n8k16\",\"warps_per_cta\":%d,\"tflops\":%.3f}\n", warps, fl / ms * 1e-9);
  }
  // HBM write bandwidth
  {
    size_t n = (size_t)1 << 29;  // 4 GiB of doubles
    double* p; CK(cudaMalloc(&p, n * 8));
    float ms = time_ms([&] { fill_kernel<<<nsm * 8, 512>>>(p, n, 1.0); }, 5);
    printf("{\"test\":\"hbm_write\",\"gbs\":%.1f}\n", n * 8.0 / ms * 1e-6);
    cudaFree(p);
  }
  // cuBLAS DGEMM
  cublasHandle_t h; cublasCreate(&h);
  for (int n : {2048, 4096, 8192}) {
    double *A, *B, *C; size_t sz = (size_t)n * n * 8;
    CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&B, sz)); CK(cudaMalloc(&C, sz));
    fill_kernel<<<nsm * 8, 512>>>(A, (size_t)n * n, 0.5); fill_kernel<<<nsm * 8, 512>>>(B, (size_t)n * n, 0.25);
    CK(cudaMemset(C, 0, sz));
    double one = 1.0, zero = 0.0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
    printf("{\"test\":\"cublas_dgemm\",\"n\":%d,\"tflops\":%.3f,\"ms\":%.3f}\n", n, 2.0 * n * (double)n * n / ms * 1e-9, ms);
    // sustained: back-to-back for ~2 s
    if (n == 8192) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int reps = (int)(2000.0f / ms) + 1;
      cudaEventRecord(e0);
      for (int r = 0; r < reps; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float tot; cudaEventElapsedTime(&tot, e0, e1);
      printf("{\"test\":\"cublas_dgemm_sustained\",\"n\":%d,\"tflops\":%.3f,\"reps\":%d}\n", n, 2.0 * n * (double)n * n * reps / tot * 1e-9, reps);
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  // cuSOLVER potrf: the library bar for the factorisation
  cusolverDnHandle_t sh; cusolverDnCreate(&sh);
  for (int n : {1000, 2000, 4096, 8192}) {
    size_t sz = (size_t)n * n * 8;
    std::vector<double> hA((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) hA[(size_t)i * n + j] = (i == j) ? (double)n : 1.0 / (1.0 + abs(i - j));
    double *A, *A0; CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&A0, sz));
    CK(cudaMemcpy(A0, hA.data(), sz, cudaMemcpyHostToDevice));
    int lwork = 0; cusolverDnDpotrf_bufferSize(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork);
    double* work; CK(cudaMalloc(&work, sizeof(double) * lwork)); int* info; CK(cudaMalloc(&info, 4));
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
      CK(cudaMemcpy(A, A0, sz, cudaMemcpyDeviceToDevice));
      float ms = 0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); cusolverDnDpotrf(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    printf("{\"test\":\"cusolver_dpotrf\",\"n\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", n, best, (double)n * n * n / 3.0 / best * 1e-9);
    cudaFree(A); cudaFree(A0); cudaFree(work); cudaFree(info);
  }
  {
    int n = 1000, B = 64; size_t sz = (size_t)n * n * 8;
    std::vector<double> hA((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) hA[(size_t)i * n + j] = (i == j) ? (double)n : 1.0 / (1.0 + abs(i - j));
    double* Aall; CK(cudaMalloc(&Aall, sz * B));
    std::vector<double*> ptrs(B); for (int b = 0; b < B; b++) ptrs[b] = Aall + (size_t)b * n * n;
    double** dptrs; CK(cudaMalloc(&dptrs, sizeof(double*) * B)); CK(cudaMemcpy(dptrs, ptrs.data(), sizeof(double*) * B, cudaMemcpyHostToDevice));
    int* info; CK(cudaMalloc(&info, 4 * B));
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      for (int b = 0; b < B; b++) CK(cudaMemcpy(ptrs[b], hA.data(), sz, cudaMemcpyHostToDevice));
      float ms = 0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); cusolverDnDpotrfBatched(sh, CUBLAS_FILL_MODE_LOWER, n, dptrs, n, info, B); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    printf("{\"test\":\"cusolver_dpotrfBatched\",\"n\":%d,\"batch\":%d,\"ms\":%.3f,\"tflops\":%.3f}\n", n, B, best, (double)B * n * n * n / 3.0 / best * 1e-9);
  }
  return 0;
}

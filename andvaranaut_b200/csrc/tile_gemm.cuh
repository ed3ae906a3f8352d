// FP64 tile GEMM primitive for sm_100a: cp.async multi-stage shared-memory pipeline feeding
// DMMA (mma.sync.aligned.m8n8k4 f64 -- the only FP64 tensor path on Blackwell; tcgen05/TMEM has no f64 kind).
//
// One CTA computes a BM x BN accumulator tile   acc(m,n) += sum_k A(m,k) * B(n,k)
// over a contiguous k range.  Operands live in global memory in one of two orientations:
//   K-contiguous ("KC"):  element (r,k) at ptr + r*ld + k      (rows of L / T used as-is)
//   R-contiguous ("RC"):  element (r,k) at ptr + k*ld + r      (columns of T: T^T T products, T*K_xs)
// Tiles are staged in shared memory in their global orientation with a +4 double pad, which makes
// both fragment access patterns bank-conflict free for 64-bit loads (ld % 16 == 4).
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA, "Matrix Fragments for mma.m8n8k4 with .f64"):
//   g = lane >> 2, t = lane & 3;  A: (row g, k t);  B: (k t, col g);  C: (row g, cols 2t, 2t+1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace avn {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

constexpr int SPAD = 4;

// Shared-memory footprint (in doubles) of one pipeline stage of one operand.
template <int ROWS, int BK, bool RC>
struct OperandTile {
  static constexpr int LD = RC ? (ROWS + SPAD) : (BK + SPAD);
  static constexpr int SIZE = RC ? (BK * LD) : (ROWS * LD);
};

// Copy one operand tile (ROWS x BK) global -> shared with 16-byte cp.async.
// row_limit: rows >= row_limit are clamped to row_limit-1 (ragged last tile; their results are discarded).
template <int ROWS, int BK, bool RC, int NTHREADS>
__device__ __forceinline__ void load_operand(double* s, const double* __restrict__ g, int64_t ld,
                                             int row_limit, int tid) {
  using OT = OperandTile<ROWS, BK, RC>;
  if (!RC) {
    constexpr int CPR = BK / 2;  // 16-byte chunks per row
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NTHREADS) {
      int r = c / CPR, kc = (c % CPR) * 2;
      int rr = r < row_limit ? r : row_limit - 1;
      cp_async16(s + r * OT::LD + kc, g + (int64_t)rr * ld + kc);
    }
  } else {
    constexpr int CPR = ROWS / 2;
    constexpr int TOTAL = BK * CPR;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NTHREADS) {
      int k = c / CPR, rc = (c % CPR) * 2;
      // ragged tiles in the R-contiguous direction are not used by any caller (npad is a tile multiple)
      cp_async16(s + k * OT::LD + rc, g + (int64_t)k * ld + rc);
    }
  }
}

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, bool A_RC_, bool B_RC_>
struct TileGemm {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr bool A_RC = A_RC_, B_RC = B_RC_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int MI = WM / 8, NI = WN / 8;
  using TA = OperandTile<BM, BK, A_RC>;
  using TB = OperandTile<BN, BK, B_RC>;
  static constexpr int STAGE_DOUBLES = TA::SIZE + TB::SIZE;
  static constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES;
  static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES * sizeof(double);

  double acc[MI][NI][2];
  // warp-id swizzle (0 or 1): which hardware warps (= SM sub-partitions) play the wm = 0 / wm = 1 roles.  Callers whose
  // wm = 1 warps skip work (triangular first slabs) alternate it between CTAs so that no sub-partition idles.
  int swz = 0;

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  // multiply-accumulate one staged k-slab (BK deep) from shared memory; only the first MACT 8-row fragments of
  // this warp are computed (MACT < MI: last block row of a matrix whose size is not a tile multiple -- the
  // other fragments are padding and keep their accumulators).
  template <int MACT>
  __device__ __forceinline__ void compute_stage(const double* sA, const double* sB, int wm,
                                                int wn, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[MACT], b[NI];
#pragma unroll
      for (int i = 0; i < MACT; i++) {
        int m = wm * WM + i * 8 + g;
        a[i] = A_RC ? sA[(kk + t) * TA::LD + m] : sA[m * TA::LD + kk + t];
      }
#pragma unroll
      for (int j = 0; j < NI; j++) {
        int n = wn * WN + j * 8 + g;
        b[j] = B_RC ? sB[(kk + t) * TB::LD + n] : sB[n * TB::LD + kk + t];
      }
#pragma unroll
      for (int i = 0; i < MACT; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }

  // acc += A[., k0:k0+klen) * B[., k0:k0+klen)^T ; klen must be a multiple of BK.
  // Apt/Bpt point at (row 0, k = 0) of the respective tiles.  All threads of the CTA must call.
  // Whole BM x BK / BN x BK tiles are always loaded (every caller works on tile-padded storage); a_rows < BM only
  // says that the 8-row fragments from a_rows (rounded up) on are padding: they issue no DMMA.
  __device__ __forceinline__ void run(double* smem, const double* Apt, int64_t lda,
                                      int a_rows, const double* Bpt, int64_t ldb, int b_rows, int klen) {
    run(smem, Apt, lda, a_rows, Bpt, ldb, b_rows, klen, [](int) {});
  }

  // As above; pre_issue(kt) is called by every thread (uniformly) before the loads of k-slab kt are issued,
  // so that a dataflow kernel can wait for operands that other CTAs are still producing.
  // skip_kt: this warp issues no DMMA for the first skip_kt k-slabs (its part of the operands is known to be
  // zero there: triangular diagonal blocks); it still takes part in the loads and barriers.
  template <typename PreIssue>
  __device__ __forceinline__ void run(double* smem, const double* Apt, int64_t lda, int a_rows, const double* Bpt,
                                      int64_t ldb, int b_rows, int klen, PreIssue pre_issue, int skip_kt = 0) {
    run(smem, Apt, lda, a_rows, Bpt, ldb, b_rows, klen, pre_issue, skip_kt, [](double*) {});
  }

  // As above; tail_issue(buf) is called once by every thread (uniformly) in the first iteration that has no k-slab left
  // to load: `buf` is the stage buffer that just became free (STAGE_DOUBLES doubles).  The caller may start cp.async
  // copies of whatever its epilogue needs into it -- they travel while the last STAGES - 1 slabs are multiplied and are
  // complete when run() returns (it ends with wait_group 0 and a barrier).  Not called when klen < BK.
  template <typename PreIssue, typename TailIssue>
  __device__ __forceinline__ void run(double* smem, const double* Apt, int64_t lda, int a_rows, const double* Bpt,
                                      int64_t ldb, int b_rows, int klen, PreIssue pre_issue, int skip_kt,
                                      TailIssue tail_issue) {
    const int tid = threadIdx.x;
    const int warp = (tid >> 5) ^ swz, lane = tid & 31;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int g = lane >> 2, t = lane & 3;
    const int KT = klen / BK;
    auto issue = [&](int kt) {
      double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
      double* sB = sA + TA::SIZE;
      const int64_t koff = (int64_t)kt * BK;
      load_operand<BM, BK, A_RC, NTHREADS>(sA, A_RC ? Apt + koff * lda : Apt + koff, lda, BM, tid);
      load_operand<BN, BK, B_RC, NTHREADS>(sB, B_RC ? Bpt + koff * ldb : Bpt + koff, ldb, BN, tid);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
      if (s < KT) {
        pre_issue(s);
        issue(s);
      }
      cp_async_commit();
    }
    // mode 0: whole tile; 1: ragged -- this warp computes its first MI/2 fragments; 2: ragged -- all of them;
    // 3: ragged -- none (the warp only moves data).  Every mode runs the same barrier sequence.
    auto mainloop = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nk = kt + STAGES - 1;
        if (nk < KT) {
          pre_issue(nk);
          issue(nk);
        } else if (nk == KT) {
          tail_issue(smem + (nk % STAGES) * STAGE_DOUBLES);
        }
        cp_async_commit();
        const double* sA = smem + (kt % STAGES) * STAGE_DOUBLES;
        if (kt >= skip_kt) {
          if constexpr (MODE == 0 || MODE == 2) compute_stage<MI>(sA, sA + TA::SIZE, wm, wn, g, t);
          if constexpr (MODE == 1) compute_stage<(MI + 1) / 2>(sA, sA + TA::SIZE, wm, wn, g, t);
        }
      }
    };
    if (a_rows >= BM) {
      mainloop(std::integral_constant<int, 0>{});
    } else {
      const int lim = (a_rows - wm * WM + 7) >> 3;   // valid 8-row fragments of this warp
      if (lim <= 0) mainloop(std::integral_constant<int, 3>{});
      else if (lim <= (MI + 1) / 2) mainloop(std::integral_constant<int, 1>{});
      else mainloop(std::integral_constant<int, 2>{});
    }
    cp_async_wait<0>();
    __syncthreads();
  }

  // Visit every accumulator element: f(row_in_tile, col_in_tile, value&)
  template <typename F>
  __device__ __forceinline__ void for_each(F f) {
    const int warp = (threadIdx.x >> 5) ^ swz, lane = threadIdx.x & 31;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) {
        int r = wm * WM + i * 8 + g, c = wn * WN + j * 8 + 2 * t;
        f(r, c, acc[i][j][0]);
        f(r, c + 1, acc[i][j][1]);
      }
  }
};

}  // namespace avn

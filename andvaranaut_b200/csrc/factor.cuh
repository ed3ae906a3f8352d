// K2: batched FP64 Cholesky L = chol(K) and triangular inverse T = L^-1 as ONE persistent dataflow kernel.
//
// The factorisation of every sample is cut into 64 x 64 tile tasks of three kinds, for block column / row k:
//   D(b,k)    A[k,k] -= sum_{m<k} L[k,m] L[k,m]^T ;  L[k,k] = chol(A[k,k]) ;  T[k,k] = L[k,k]^-1 ;  sum log L_ii
//   P(b,k,i)  L[i,k]  = (A[i,k] - sum_{m<k} L[i,m] L[k,m]^T) T[k,k]^T                          (i > k, left-looking)
//   R(b,k,j)  T[k,j]  = -T[k,k] sum_{m=j}^{k-1} L[k,m] T[m,j]                                  (j < k, row k of L^-1)
// Tasks are numbered step-major, the sample index fastest: first D(.,0); then for every step s the tiles
// P(.,s,s+1), `dgap` more slots of P/R work, the look-ahead diagonal task D(.,s+1) (it only needs P(.,s,s+1), and
// by the time step s+1 starts it has long finished), the remaining P(.,s,i) by rising i and R(.,s,j) by rising j
// (longest first).  CTAs are persistent: each takes the next task number from a global ticket counter, waits on
// per-(sample, block) progress flags for its operands -- slab by slab inside the k loop, so that only the last
// 64-deep slab of a tile product is on the critical path -- runs the DMMA tile product and publishes its own
// flag.  A task only ever waits on tasks with smaller numbers, and a number is handed out only to a CTA that is
// already running, so the waits cannot deadlock; with B * nb tasks per step and ~450 resident CTAs they are
// rarely taken at all.  There are no launch boundaries, no wave quantisation and no host round trips inside the
// factorisation (the previous version needed 3 launches per block column).
//
// Flags (int32, zeroed by the host before the launch):
//   lflag[b][i] = number of final 64-wide block columns in block row i of L
//   tflag[b][j] = number of final block rows in block column j of T, counted as (row index + 1): tflag[k] = k + 1, published
//                 by the diagonal task, says that T[k,k] is there -- what the panel and inverse tasks of step k wait for
#pragma once
#include "avn_dev.cuh"
#include "tile_gemm.cuh"

namespace avn {

// Pipeline shape per launch kind.  Throughput launches: 32-deep slabs x 2 stages (72 KB, still 3 CTAs per SM) -- half the
// per-slab barriers of 16 x 3; measured at HEAD of round 2: factor 11.88 -> 11.69 ms at c2 / B=64, 13.61 -> 13.48 ms at
// c3 (the plain tile GEMM: 34.4 -> 35.2 TF).  Chain launches (few samples) keep 16 x 3: the tasks next to the chain wait
// slab by slab, and finer slabs let them start sooner (factor 0.584 ms against 0.621 ms at B = 1).
#ifndef AVN_FAC_BK
#define AVN_FAC_BK 32
#define AVN_FAC_STAGES 2
#endif
#ifndef AVN_POLL_NS
#define AVN_POLL_NS 64      // pause between two polls of a progress flag (measured: see DESIGN.md)
#endif
constexpr int FAC_THREADS = 128;
constexpr int FAC_LDS = TILE + SPAD;                                 // 68: conflict-free fragment loads both ways
constexpr size_t cmax(size_t a, size_t b) { return a > b ? a : b; }
template <bool FUSED>
struct FacCfg {
  static constexpr int BK = FUSED ? 16 : AVN_FAC_BK, STAGES = FUSED ? 3 : AVN_FAC_STAGES;
  static constexpr int SPB = TILE / BK;   // pipeline slabs per 64-deep block
  using KK = TileGemm<64, 64, BK, 32, 32, STAGES, false, false>;
  using KR = TileGemm<64, 64, BK, 32, 32, STAGES, false, true>;
  // two staged 64 x 64 tiles for the epilogues alias the pipeline buffers
  static constexpr size_t SMEM2 = cmax((size_t)2 * TILE * FAC_LDS * 8, cmax(KK::SMEM_BYTES, KR::SMEM_BYTES));
};
constexpr size_t FAC_SMEM_BYTES = FacCfg<false>::SMEM2;
// fused-panel launches (few samples): a third staged tile behind the first two and behind the pipeline buffers
constexpr int FAC_THIRD_OFF = (int)(FacCfg<true>::SMEM2 / 8);   // in doubles
constexpr size_t FAC_SMEM_BYTES_FUSED = FacCfg<true>::SMEM2 + (size_t)TILE * FAC_LDS * 8;

#ifdef AVN_FACTOR_PROF
// timeline of the critical chain (B = 1): thread 0 logs (task type, k, event, globaltimer ns) for the diagonal tasks
// and the panel tile right below the diagonal into prof[16 ...] (4 long long per entry, entry count in prof[15])
__device__ __forceinline__ void fprof_event(long long* prof, int type, int k, int ev) {
  if (threadIdx.x != 0 || !prof) return;
  unsigned long long ns;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
  const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(prof) + 15, 1ull);
  if (slot < 4000) {
    long long* e = prof + 16 + 4 * slot;
    e[0] = type; e[1] = k; e[2] = ev; e[3] = (long long)ns;
  }
}
#define FEVENT(type, k, ev) fprof_event(fa.prof, type, k, ev)
#define FPROF_DECL long long fp_t0 = clock64(), fp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define FPROF(slot)                                   \
  do {                                                \
    long long fp_t1 = clock64();                      \
    fp_acc[slot] += fp_t1 - fp_t0;                    \
    fp_t0 = fp_t1;                                    \
  } while (0)
#define FPROF_FLUSH(p)                                                                       \
  do {                                                                                       \
    if (threadIdx.x == 0 && (p))                                                             \
      for (int q = 0; q < 8; q++) atomicAdd(reinterpret_cast<unsigned long long*>(p) + q, (unsigned long long)fp_acc[q]); \
  } while (0)
#else
#define FEVENT(type, k, ev)
#define FPROF_DECL
#define FPROF(slot)
#define FPROF_FLUSH(p)
#endif

struct FactorArgs {
  double* L;            // [B][npad][npad]  K on entry (lower block triangle), L on exit
  double* T;            // [B][npad][npad]  T = L^-1 (lower block triangle; diagonal blocks with explicit zeros)
  double* fpart;        // [B][nb][2]       [.][1] = sum of log L_ii over the block (slot 0 belongs to beta_reduce_kernel)
  const double* z;      // [B][npad]        converted outputs: beta = T z is formed tile by tile as T is produced
  double* beta;         // [B][npad]        receives the diagonal-block share T_kk z_k (beta_reduce_kernel adds the rest)
  int32_t* info;        // [B]              first non-positive pivot (1-based) or 0
  int32_t* lflag;       // [B][nb]
  int32_t* tflag;       // [B][nb]
  int32_t* sflag;       // [B][nb]  fused-panel launches: 1 = S of the tile (i, i-1) is parked in the T slab
  int32_t* dflag;       // [B][nb]  fused-panel launches: 1 = Dpre(b,i) has left the partial diagonal block in place of A_ii
  int32_t* ctl;         // [0] ticket counter, [1] abort flag (a wait exceeded its bound)
  int npad, nb, B;
  int n;                // valid rows (N <= npad): tiles of the last block row skip their padding fragments
  int want_inverse;     // 0: skip the R tasks (log-likelihood only)
  int dgap;             // slots between P(.,s,s+1) and the look-ahead D(.,s+1)
  unsigned max_spins;   // bound of every flag wait in polls (avn_gp_set_debug; default 2^26, several seconds)
  int fault;            // fault injection (tests): 1 = the panel task P(0,0,1) never publishes its flag
  int fuse_panel;       // 1: diagonal tasks form their own panel tile L[k,k-1] (few samples; needs FAC_SMEM_BYTES_FUSED)
  long long* prof;      // AVN_FACTOR_PROF builds: 8 cycle counters (ticket, wait, gemm, wait T_kk, epilogue, diag, -, -)
};

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// thread 0 polls until *flag >= need.  Bounded: after max_spins polls the abort flag ctl[1] is raised, which ends every
// later wait at once; finalize_kernel / abort_check_kernel turn it into info = -1 (results invalid), so a timed-out wait
// is reported, never returned as data.
__device__ __forceinline__ void wait_flag(const int32_t* flag, int need, int32_t* ctl, unsigned max_spins) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire(flag) < need) {
      __nanosleep(AVN_POLL_NS);
      if ((++spins & 1023u) == 0) {
        if (ld_acquire(ctl + 1) != 0) break;
        if (spins > max_spins) {
          atomicExch(ctl + 1, 1);
          break;
        }
      }
    }
  }
  __syncthreads();
}

// Slab-wise operand wait inside TileGemm::run: k-slab kt belongs to the 64-deep block m = m0 + kt / spb, which
// needs both progress flags >= m + 1.  `known` caches the smaller flag value seen last, so the flags are only
// read again (thread 0, then one barrier) when the product runs ahead of what was known to be final.
struct SlabWaiter {
  const int32_t* fa;
  const int32_t* fb;
  int32_t* ctl;
  int* s_known;
  int m0, known;
  unsigned max_spins;
  const int32_t* fc = nullptr;   // optional third flag with its own threshold (the chain's combined wait)
  int need_c = 0;
  int spb = 1;                   // pipeline slabs per 64-deep block (FacCfg::SPB)
  __device__ __forceinline__ void operator()(int kt) {
    if (kt % spb) return;
    const int need = m0 + kt / spb + 1;
    if (need <= known) return;
    if (threadIdx.x == 0) {
      unsigned spins = 0;
      int have;
      for (;;) {
        const int va = ld_acquire(fa), vb = ld_acquire(fb);
        have = va < vb ? va : vb;
        if (fc && ld_acquire(fc) < need_c) have = 0;
        if (have >= need) break;
        __nanosleep(AVN_POLL_NS);
        if ((++spins & 1023u) == 0) {
          if (ld_acquire(ctl + 1) != 0) { have = 0x7fffffff; break; }
          if (spins > max_spins) {
            atomicExch(ctl + 1, 1);
            have = 0x7fffffff;
            break;
          }
        }
      }
      *s_known = have;
    }
    __syncthreads();
    known = *s_known;
  }
};

// All threads have written their part of a tile: publish the flag.  The barrier orders the CTA's stores before
// One release store by thread 0 after a CTA barrier publishes the whole CTA's tile: st.release.gpu is cumulative, the
// barrier orders the other threads' stores before it (PTX memory model: bar.sync synchronises the CTA, release makes
// everything that happens-before it visible to the acquirer).  No separate __threadfence(): that is a second,
// sequentially-consistent MEMBAR.SC.GPU plus an L1 invalidation (CCTL.IVALL) in front of the MEMBAR.ALL.GPU the
// release store already carries -- measured ~4.5 us per publish under load at B = 64.  Consumers read the flag with
// ld.acquire in one thread, pass a barrier and read the data through L2 (cp.async.cg / ld.global.cg).
__device__ __forceinline__ void publish(int32_t* flag, int value) {
  __syncthreads();
  if (threadIdx.x == 0) st_release(flag, value);
}
__device__ __forceinline__ void publish2(int32_t* f0, int v0, int32_t* f1, int v1) {
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release(f0, v0);
    st_release(f1, v1);
  }
}

// accumulators <- minus a 64 x 64 global tile (the product then accumulates L L^T - A = -X on top of it)
__device__ __forceinline__ void load_neg_tile(double (&acc)[4][4][2], const double* __restrict__ g, int64_t ld, int wm,
                                              int wn, int gq, int t) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
      const double2 v = __ldcg(reinterpret_cast<const double2*>(g + (int64_t)r * ld + c));
      acc[i][j][0] = -v.x;
      acc[i][j][1] = -v.y;
    }
}

// acc(m,n) += sum_c sA[m][c] * (B_KN ? sB[c][n] : sB[n][c]); 64x64x64 from shared memory, warp tile 32x32
template <bool B_KN>
__device__ __forceinline__ void smem_gemm64x(double (&acc)[4][4][2], const double* sA, const double* sB, int wm, int wn,
                                             int g, int t, int kmax) {
#pragma unroll 4
  for (int kk = 0; kk < kmax; kk += 4) {
    double a[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = sA[(wm * 32 + i * 8 + g) * FAC_LDS + kk + t];
#pragma unroll
    for (int j = 0; j < 4; j++)
      bb[j] = B_KN ? sB[(kk + t) * FAC_LDS + wn * 32 + j * 8 + g] : sB[(wn * 32 + j * 8 + g) * FAC_LDS + kk + t];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
  }
}

// copy a 64 x 64 global tile (row stride ld) into a staged shared tile, bypassing L1 (the data was written by
// another SM during this launch)
__device__ __forceinline__ void stage_tile(double* s, const double* __restrict__ g, int64_t ld) {
  constexpr int PER = TILE * TILE / 2 / FAC_THREADS;   // 16 double2 per thread: all loads issued before the first store
  double2 v[PER];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int e = threadIdx.x + q * FAC_THREADS, r = e >> 5, c = (e & 31) * 2;
    v[q] = __ldcg(reinterpret_cast<const double2*>(g + (int64_t)r * ld + c));
  }
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int e = threadIdx.x + q * FAC_THREADS, r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(&s[r * FAC_LDS + c]) = v[q];
  }
}

// the same copy by cp.async (L2 -> shared memory, no registers): the caller commits the group and waits for it later
__device__ __forceinline__ void stage_tile_async(double* s, const double* __restrict__ g, int64_t ld) {
#pragma unroll
  for (int q = 0; q < TILE * TILE / 2 / FAC_THREADS; q++) {
    const int e = threadIdx.x + q * FAC_THREADS, r = e >> 5, c = (e & 31) * 2;
    cp_async16(s + r * FAC_LDS + c, g + (int64_t)r * ld + c);
  }
}

// ------------------------------------------------------------------------------------------------
// 64 x 64 diagonal block: Cholesky, then the triangular inverse; 128 threads.
// Cholesky in four 16-wide panels:
//   (1) ONE warp factors the 16 x 16 diagonal sub-block in registers, one row per lane; pivot and column j of L travel
//       by warp shuffles -- a warp-shuffle panel factorisation with no barrier inside its 16 dependent steps; the step
//       is one basic block (branch-free pivot test and rsqrt), so the scheduler overlaps the rank-1 update of step j
//       with the pivot chain (shuffle, rsqrt, scale) of step j + 1;
//   (2) rows below:  x L_pp^T = a  by substitution, one row per thread in registers, L_pp broadcast from shared memory;
//   (3) rank-16 update of the trailing sub-matrix by DMMA on 8 x 8 fragments of its lower triangle, with look-ahead:
//       only the next diagonal sub-block is updated before the next panel's factorisation starts; the rest of the
//       update runs beside it on warps 2 and 3.
// Inverse T = L^-1 in 16 x 16 blocks, T_pp = L_pp^-1 (one column per lane) and T_ij = -T_ii sum_{m=j..i-1} L_im T_mj by
// DMMA.  Block rows 0..2 are produced by warp 1 WHILE warp 0 factors the next panel (the other warps idle there anyway;
// everything that job reads is final, what it writes nobody else touches); only block row 3 is left for a short tail
// (warp 0 inverts L_33 while warps 1..3 form the three W_3j, then T_3j = -T_33 W_3j).
// History (instrumented build, cycles per diagonal task at B = 1, where these tasks ARE the critical path): a
// register-resident column sweep of the whole block with one __syncthreads per column 85 k; panels with the inverse
// fused into the warp elimination 52 k; Cholesky first + parallel sub-block inverses 36 k; branch-free steps 30.5 k;
// inverse overlapped with the factorisation 27 k; look-ahead trailing update 24 k.
// sA holds A (lower part) on entry and L (zeros above the diagonal) on exit; sT receives T (zeros above the
// diagonal); dval[j] = L_jj; sinv[j] = 1 / L_jj; *s_bad = first non-positive pivot (1-based, global index).
// ------------------------------------------------------------------------------------------------
// ---- warp-level 16 x 16 block operations on the staged tiles (row pitch FAC_LDS), used by the inverse half ----
// inverse of the lower-triangular sub-block at (c0, c0) of sA into sT: lane c owns column c (forward substitution down
// the column, two interleaved partial sums); T_cc = sinv[c] = rsqrt(pivot).  One warp; lanes 16..31 shadow 0..15.
__device__ __forceinline__ void warp_inv16(const double* sA, double* sT, const double* sinv, int c0, int lane) {
  constexpr int PB = 16;
  const int c = lane & 15;
  double tcol[PB];
#pragma unroll
  for (int r = 0; r < PB; r++) {
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int m = 0; m < r; m++) {
      const double tm = (m >= c) ? tcol[m] : 0.0;
      if (m & 1) acc1 = fma(sA[(c0 + r) * FAC_LDS + c0 + m], tm, acc1);
      else acc0 = fma(sA[(c0 + r) * FAC_LDS + c0 + m], tm, acc0);
    }
    const double ir = sinv[c0 + r];
    tcol[r] = (r == c) ? ir : ((r > c) ? -(acc0 + acc1) * ir : 0.0);
  }
  if (lane < PB) {
#pragma unroll
    for (int r = 0; r < PB; r++) sT[(c0 + r) * FAC_LDS + c0 + c] = tcol[r];
  }
}
// acc (2 x 2 fragments of 8 x 8) += A(16 x 16 at A) * B(16 x 16 at B), both row-major with pitch FAC_LDS.  One warp.
__device__ __forceinline__ void warp_mma16(double (&acc)[2][2][2], const double* A, const double* B, int gq, int t) {
#pragma unroll
  for (int kk = 0; kk < 16; kk += 4)
#pragma unroll
    for (int fi = 0; fi < 2; fi++) {
      const double a = A[(fi * 8 + gq) * FAC_LDS + kk + t];
#pragma unroll
      for (int fj = 0; fj < 2; fj++) dmma884(acc[fi][fj][0], acc[fi][fj][1], a, B[(kk + t) * FAC_LDS + fj * 8 + gq]);
    }
}
__device__ __forceinline__ void warp_store16(double* C, const double (&acc)[2][2][2], double sign, int gq, int t) {
#pragma unroll
  for (int fi = 0; fi < 2; fi++)
#pragma unroll
    for (int fj = 0; fj < 2; fj++)
      *reinterpret_cast<double2*>(&C[(fi * 8 + gq) * FAC_LDS + fj * 8 + 2 * t]) =
          make_double2(sign * acc[fi][fj][0], sign * acc[fi][fj][1]);
}
// T_ij = -T_ii sum_{m=j..i-1} L_im T_mj for one off-diagonal 16 x 16 block, by ONE warp (W parked in the destination)
__device__ __forceinline__ void warp_toff16(const double* sA, double* sT, int bi, int bj, int gq, int t) {
  constexpr int PB = 16;
  double w[2][2][2] = {};
  for (int m = bj; m < bi; m++) warp_mma16(w, sA + (bi * PB) * FAC_LDS + m * PB, sT + (m * PB) * FAC_LDS + bj * PB, gq, t);
  double* dst = sT + (bi * PB) * FAC_LDS + bj * PB;
  warp_store16(dst, w, 1.0, gq, t);
  __syncwarp();
  double x[2][2][2] = {};
  warp_mma16(x, sT + (bi * PB) * FAC_LDS + bi * PB, dst, gq, t);
  __syncwarp();
  warp_store16(dst, x, -1.0, gq, t);
  __syncwarp();
}

// trailing update of panel pp (columns c0 = 16 pp ..): A_rc -= sum_m L_r,c0+m L_c,c0+m on 8 x 8 fragments (fi >= fj) of the
// lower triangle of the trailing block, fragment pairs numbered row by row; this warp takes pairs q0, q0 + qstep, ... < q1.
// Pairs 0, 1, 2 are the next diagonal 16 x 16 sub-block.  Conflict-free fragment loads (FAC_LDS).
__device__ __forceinline__ void syrk_pairs(double* sA, int pp, int q0, int q1, int qstep, int gq, int t) {
  constexpr int PB = 16;
  const int c0 = PB * pp, f0 = (c0 + PB) / 8;
  for (int q = q0; q < q1; q += qstep) {
    int fi = 0;
    while ((fi + 1) * (fi + 2) / 2 <= q) fi++;
    const int fj = q - fi * (fi + 1) / 2;
    const int r0 = 8 * (f0 + fi), c = 8 * (f0 + fj);
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int kk = 0; kk < PB; kk += 4)
      dmma884(acc0, acc1, sA[(r0 + gq) * FAC_LDS + c0 + kk + t], sA[(c + gq) * FAC_LDS + c0 + kk + t]);
    double2* dst = reinterpret_cast<double2*>(&sA[(r0 + gq) * FAC_LDS + c + 2 * t]);
    double2 cur = *dst;
    cur.x -= acc0;
    cur.y -= acc1;
    *dst = cur;
  }
}

// 1 / sqrt(p) for a normal positive p: the hardware estimate (MUFU.RSQ64H, ~2^-22) and one third-order correction
// y (1 + e/2 + 3 e^2 / 8), e = 1 - p y^2 -- the fast path of the library rsqrt() without its range check and call, so
// that a whole pivot step stays one basic block and the scheduler can overlap it with the updates of the previous one.
__device__ __forceinline__ double rsqrt_nobranch(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  const double e = fma(-(y * y), p, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(c, y * e, y);
}

#ifdef AVN_FACTOR_PROF
#define DPROF(slot)                                                                                  \
  do {                                                                                               \
    if (threadIdx.x == 0 && dprof) {                                                                 \
      const long long dp_t1 = clock64();                                                             \
      atomicAdd(reinterpret_cast<unsigned long long*>(dprof) + (slot), (unsigned long long)(dp_t1 - dp_t0)); \
      dp_t0 = dp_t1;                                                                                 \
    }                                                                                                \
  } while (0)
#else
#define DPROF(slot)
#endif
__device__ __forceinline__ void diag_chol_inv_blocked(double* sA, double* sT, double* dval /*[64]*/,
                                                      double* sinv /*[64]*/, int* s_bad, int pivot_base,
                                                      long long* dprof = nullptr) {
  constexpr int PB = 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, t = lane & 3;
#ifdef AVN_FACTOR_PROF
  long long dp_t0 = clock64();
#endif
  for (int p = 0; p < TILE / PB; p++) {
    const int c0 = PB * p, c1 = c0 + PB;
    if (warp == 0) {
      // (1) Cholesky of the 16 x 16 diagonal sub-block, one row per lane in registers (lanes 16..31 shadow lanes
      // 0..15 so that every shuffle is a full-warp one).  Pivot and column j of L travel by shuffles: no barrier,
      // nothing in shared memory inside the 16 dependent steps.
      const int i = lane & 15;
      double x[PB];
#pragma unroll
      for (int c = 0; c < PB; c++) x[c] = sA[(c0 + i) * FAC_LDS + c0 + c];
      double myinv = 1.0;
      int badj = 0;       // first non-positive pivot of this sub-block (1-based), 0: none
      // Software-pipelined by hand: the pivot of step j + 1 (lane j+1's own A_jj, complete as soon as that lane has
      // squared its entry of column j) is broadcast and its rsqrt started BEFORE the rest of column j is exchanged and
      // applied -- a warp issues in order, so with the exchange first the ~100-cycle shuffle + rsqrt chain of the next
      // step only began after the exchange's store -> barrier -> load -> FMA sequence (~200 cycles per step instead of
      // the chain's ~120).
      double piv = __shfl_sync(0xffffffffu, x[0], 0);
      bool bad = !(piv > 0.0);   // also catches NaN
      badj = bad ? 1 : 0;
      piv = bad ? 1.0 : piv;
      double inv = rsqrt_nobranch(piv);
#pragma unroll
      for (int j = 0; j < PB; j++) {
        myinv = (i == j) ? inv : myinv;
        const double li = (i == j) ? piv * inv : x[j] * inv;        // column j of L (rows >= j)
        if (j + 1 < PB) {
          // the next pivot needs only lane j+1's own square
          const double pv = fma(-li, li, x[j + 1]);
          piv = __shfl_sync(0xffffffffu, pv, j + 1);
          bad = !(piv > 0.0);
          badj = (bad && badj == 0) ? j + 2 : badj;
          piv = bad ? 1.0 : piv;
          inv = rsqrt_nobranch(piv);
          // column j of L reaches the other lanes through 16 doubles of shared memory (one store, then 128-bit broadcast
          // loads) instead of one 64-bit shuffle per entry: 15 - j shuffles = 2 (15 - j) SHFL at a quarter-rate issue
          // slot each made the step issue-bound.  Two buffers (the panel's own, still unused slots of dval / sinv)
          // alternate, so a step's store cannot overtake the previous step's loads.  Same operands, same FMAs as the
          // shuffle form: bit-identical.
          double* colbuf = (j & 1) ? sinv + c0 : dval + c0;
          if (lane < PB) colbuf[i] = li;
          __syncwarp();
#pragma unroll
          for (int q = (j + 1) / 2; q < PB / 2; q++) {
            const double2 lc = reinterpret_cast<const double2*>(colbuf)[q];
            if (2 * q > j) x[2 * q] = fma(-li, lc.x, x[2 * q]);      // right of the diagonal: never used
            x[2 * q + 1] = fma(-li, lc.y, x[2 * q + 1]);
          }
        }
        x[j] = li;
      }
      __syncwarp();   // the last column loads are done before dval / sinv receive their final values
      DPROF(8);
      if (lane == 0 && badj != 0 && *s_bad == 0) *s_bad = pivot_base + c0 + badj;
      if (lane < PB) {
#pragma unroll
        for (int c = 0; c < PB; c++) {
          sA[(c0 + i) * FAC_LDS + c0 + c] = (c <= i) ? x[c] : 0.0;
          if (c == i) dval[c0 + i] = x[c];
        }
        sinv[c0 + i] = myinv;       // 1 / L_ii (rsqrt of the pivot, as T_ii is defined)
      }
    } else if (warp >= 2 && p >= 1) {
      // look-ahead: the previous panel updated only the next diagonal sub-block before this panel's factorisation
      // started; the rest of its trailing update (rows below that sub-block) runs here, beside warp 0
      const int nf = TILE / 8 - 2 * p;
      syrk_pairs(sA, p - 1, 3 + (warp - 2), nf * (nf + 1) / 2, 2, gq, t);
    } else if (warp >= 2 && p == 0) {
      // zeros above the diagonal outside the diagonal sub-blocks (L and T): six 16 x 16 blocks that nothing else touches,
      // filled by the two warps that idle during the first panel
      for (int e = tid - 64; e < 6 * PB * PB; e += 64) {
        const int blk = e >> 8, rr = (e >> 4) & 15, cc = e & 15;
        // blocks (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
        const int br = blk < 3 ? 0 : (blk < 5 ? 1 : 2), bc = blk < 3 ? blk + 1 : (blk < 5 ? blk - 1 : 3);
        sA[(br * PB + rr) * FAC_LDS + bc * PB + cc] = 0.0;
        sT[(br * PB + rr) * FAC_LDS + bc * PB + cc] = 0.0;
      }
    } else if (warp == 1 && p >= 1) {
      // the inverse half trails the factorisation by one panel on a warp that would otherwise idle here: everything it
      // reads (L blocks of rows < p, sinv) is final, everything it writes (T blocks of rows < p) is touched by nobody else
      warp_inv16(sA, sT, sinv, PB * (p - 1), lane);
      __syncwarp();
      for (int bj = p - 2; bj >= 0; bj--) warp_toff16(sA, sT, p - 1, bj, gq, t);
    }
    DPROF(9);
    __syncthreads();
    if (c1 < TILE) {
      // (2) rows below the sub-block: x L_pp^T = a by substitution, one row per thread in registers; L_pp is read from
      // shared memory (every thread the same address: broadcast).  As soon as x_m is known every later column is updated
      // (independent FMAs), so the dependent depth is 16 x (multiply + FMA).
      {
        const int r = c1 + tid;
        if (r < TILE) {
          double a[PB];
#pragma unroll
          for (int m = 0; m < PB; m++) a[m] = sA[r * FAC_LDS + c0 + m];
#pragma unroll
          for (int m = 0; m < PB; m++) {
            a[m] *= sinv[c0 + m];
#pragma unroll
            for (int c = m + 1; c < PB; c++) a[c] = fma(-a[m], sA[(c0 + c) * FAC_LDS + c0 + m], a[c]);
          }
#pragma unroll
          for (int m = 0; m < PB; m++) sA[r * FAC_LDS + c0 + m] = a[m];
        }
      }
      __syncthreads();
      DPROF(10);
      // (3) trailing update, look-ahead form: only the next diagonal 16 x 16 sub-block (three fragment pairs, one warp
      // each) stands between this panel and the next factorisation; the rest follows beside it (see above)
      if (warp < 3) syrk_pairs(sA, p, warp, 3, 3, gq, t);
      __syncthreads();
      DPROF(11);
    }
  }
  // tail: only block row 3 of T is left.  warp 0 inverts L_33 while warps 1..3 form W_3j = sum_m L_3m T_mj (j = 0, 1, 2;
  // every T_mj with m < 3 was finished during the panel loop); then T_3j = -T_33 W_3j.
  __syncthreads();
  {
    constexpr int LAST = TILE / PB - 1;
    double w[2][2][2] = {};
    double* dst = sT + (LAST * PB) * FAC_LDS + (warp - 1) * PB;
    if (warp == 0) {
      warp_inv16(sA, sT, sinv, PB * LAST, lane);
    } else {
      const int bj = warp - 1;
      for (int m = bj; m < LAST; m++)
        warp_mma16(w, sA + (LAST * PB) * FAC_LDS + m * PB, sT + (m * PB) * FAC_LDS + bj * PB, gq, t);
      warp_store16(dst, w, 1.0, gq, t);
    }
    DPROF(13);
    __syncthreads();
    if (warp != 0) {
      double x[2][2][2] = {};
      warp_mma16(x, sT + (LAST * PB) * FAC_LDS + LAST * PB, dst, gq, t);
      __syncwarp();
      warp_store16(dst, x, -1.0, gq, t);
    }
  }
  __syncthreads();
  DPROF(12);
}

// blk -= X X^T on the 36 8 x 8 fragments (fi >= fj) of the lower triangle of a 64 x 64 block, X staged in shared memory
// (pitch FAC_LDS); the last-slab update of a diagonal task, on the critical path of a single factorisation.  Nine
// fragments per warp, register-blocked: warps 0..2 take the 3 x 3 rectangles rows {5,6,7} x columns {0,1,2}, rows {2,3,4}
// x columns {0,1,2}, rows {5,6,7} x columns {3,4,5}; warp 3 the three 2 x 2 triangles at 0, 3 and 6 -- six operand
// fragments feed nine DMMAs per 4-deep step (one fragment at a time needed 18 and sat at ~60 cycles per DMMA on their
// latency), eighteen independent accumulator chains.  Per fragment: two accumulator pairs taking the even / odd 4-deep
// steps, summed at the end -- the order every earlier version used, so the bits are unchanged.
__device__ __forceinline__ void diag_update_blocked(double* blk, const double* sX, int warp, int gq, int t) {
  int fr[3], fc[3];
  if (warp < 3) {
    const int r0 = (warp == 1) ? 2 : 5, c0 = (warp == 2) ? 3 : 0;
#pragma unroll
    for (int u = 0; u < 3; u++) { fr[u] = r0 + u; fc[u] = c0 + u; }
    double acc[3][3][4];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;
#pragma unroll
    for (int kk = 0; kk < TILE; kk += 8) {
      double a[3][2], b[3][2];
#pragma unroll
      for (int u = 0; u < 3; u++) {
        a[u][0] = sX[(8 * fr[u] + gq) * FAC_LDS + kk + t];
        a[u][1] = sX[(8 * fr[u] + gq) * FAC_LDS + kk + 4 + t];
        b[u][0] = sX[(8 * fc[u] + gq) * FAC_LDS + kk + t];
        b[u][1] = sX[(8 * fc[u] + gq) * FAC_LDS + kk + 4 + t];
      }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          dmma884(acc[i][j][0], acc[i][j][1], a[i][0], b[j][0]);
          dmma884(acc[i][j][2], acc[i][j][3], a[i][1], b[j][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double2* dst = reinterpret_cast<double2*>(&blk[(8 * fr[i] + gq) * FAC_LDS + 8 * fc[j] + 2 * t]);
        double2 cur = *dst;
        cur.x -= acc[i][j][0] + acc[i][j][2];
        cur.y -= acc[i][j][1] + acc[i][j][3];
        *dst = cur;
      }
  } else {
    // triangles {(f,f), (f+1,f), (f+1,f+1)}, f = 0, 3, 6: operand fragments f and f+1
    double acc[3][3][4];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;
#pragma unroll
    for (int kk = 0; kk < TILE; kk += 8) {
      double x[3][2][2];   // [triangle][fragment f / f+1][even / odd step]
#pragma unroll
      for (int u = 0; u < 3; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
          x[u][v][0] = sX[(8 * (3 * u + v) + gq) * FAC_LDS + kk + t];
          x[u][v][1] = sX[(8 * (3 * u + v) + gq) * FAC_LDS + kk + 4 + t];
        }
#pragma unroll
      for (int u = 0; u < 3; u++) {
        dmma884(acc[u][0][0], acc[u][0][1], x[u][0][0], x[u][0][0]);   // (f, f)
        dmma884(acc[u][0][2], acc[u][0][3], x[u][0][1], x[u][0][1]);
        dmma884(acc[u][1][0], acc[u][1][1], x[u][1][0], x[u][0][0]);   // (f+1, f)
        dmma884(acc[u][1][2], acc[u][1][3], x[u][1][1], x[u][0][1]);
        dmma884(acc[u][2][0], acc[u][2][1], x[u][1][0], x[u][1][0]);   // (f+1, f+1)
        dmma884(acc[u][2][2], acc[u][2][3], x[u][1][1], x[u][1][1]);
      }
    }
#pragma unroll
    for (int u = 0; u < 3; u++)
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int fi = 3 * u + (q > 0), fj = 3 * u + (q > 1);
        double2* dst = reinterpret_cast<double2*>(&blk[(8 * fi + gq) * FAC_LDS + 8 * fj + 2 * t]);
        double2 cur = *dst;
        cur.x -= acc[u][q][0] + acc[u][q][2];
        cur.y -= acc[u][q][1] + acc[u][q][3];
        *dst = cur;
      }
  }
}

// X = S T_kk^T on 8 x 8 fragments, S and T_kk staged in shared memory (pitch FAC_LDS).  T_kk is lower triangular: the
// fragment column jf needs k < 8 jf + 8 only.  Warp w takes the fragment columns w and 7 - w (18 four-deep steps per row
// fragment between them, the same for every warp) and all eight row fragments: 144 DMMAs per warp, against 128 / 256
// with the 32 x 32 quadrant layout.  xa[ii][c] = fragment (row ii, column c ? 7 - w : w).
__device__ __forceinline__ void panel_product(double (&xa)[8][2][2], const double* sS, const double* sTk, int warp, int gq,
                                              int t) {
  // (a variant that shared the S fragments between the two columns and loaded the operands of step k + 4 before the
  // DMMAs of step k was slower both in the 168-register kernel -- 3.9 us against 2.3 us per chain link, +2 % on throughput
  // launches -- and as a chain-only instantiation with 248 registers: factor 0.584 -> 0.64 ms at B = 1; a layout with
  // row fragments 2w, 2w+1 x all eight column fragments per warp -- 104 operand loads instead of 162 -- changed nothing)
#pragma unroll
  for (int ii = 0; ii < 8; ii++) xa[ii][0][0] = xa[ii][0][1] = xa[ii][1][0] = xa[ii][1][1] = 0.0;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const int jf = c ? 7 - warp : warp;
    const int kmax = 8 * jf + 8;
#pragma unroll 2
    for (int kk = 0; kk < kmax; kk += 4) {
      const double bv = sTk[(8 * jf + gq) * FAC_LDS + kk + t];
#pragma unroll
      for (int ii = 0; ii < 8; ii++) dmma884(xa[ii][c][0], xa[ii][c][1], sS[(8 * ii + gq) * FAC_LDS + kk + t], bv);
    }
  }
}

// FUSED: the fused-panel / chain launch for few samples (fa.fuse_panel == 1); its own instantiation, so that the
// throughput kernel keeps the register allocation it had before the chain code existed (160 registers, no spills)
template <bool FUSED>
__global__ void __launch_bounds__(FAC_THREADS, FUSED ? 2 : 3) factor_kernel(FactorArgs fa) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_ticket;
  __shared__ int s_known;
  __shared__ int s_bad;
  __shared__ int s_tkk;   // throughput launches: T_kk was there when the task started (its tile is prefetched, see P / R)
  __shared__ __align__(16) double s_dval[TILE];
  __shared__ __align__(16) double s_inv[TILE];
  using FacKK = typename FacCfg<FUSED>::KK;
  using FacKR = typename FacCfg<FUSED>::KR;
  constexpr int FAC_BK = FacCfg<FUSED>::BK, FAC_SPB = FacCfg<FUSED>::SPB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % 2, wn = warp / 2, gq = lane >> 2, t = lane & 3;
  const int npad = fa.npad, nb = fa.nb, B = fa.B;
  const int per_step = B * nb;
  const int total = per_step * nb;
  double* sA = smem;
  double* sB = smem + TILE * FAC_LDS;
  FPROF_DECL;
  // The ticket of the NEXT task is drawn while the current one runs (thread 0 keeps it in a register), so the round
  // trip of the atomic is off the path between two tasks.  Still deadlock-free: the smallest unfinished ticket is never
  // one that is merely held -- its holder would be running a smaller, unfinished one -- so it is always running, and
  // everything it waits on has finished.
  // Throughput launches only: with few samples the CTAs run far ahead of the chain of diagonal tasks and mostly wait, and
  // a ticket held behind a waiting task would keep the next link of that chain from starting.
  constexpr bool prefetch = !FUSED;
  int next_ticket = 0;
  if (tid == 0 && prefetch) next_ticket = atomicAdd(fa.ctl, 1);
  for (;;) {
    __syncthreads();  // everyone is done with s_ticket and the staged tiles of the previous task
    if (tid == 0) s_ticket = prefetch ? next_ticket : atomicAdd(fa.ctl, 1);
    __syncthreads();
    const int ticket = s_ticket;
    FPROF(0);
    if (ticket >= total) break;
    if (tid == 0 && prefetch) next_ticket = atomicAdd(fa.ctl, 1);
    // ticket -> task (see the header comment for the order)
    int type, k, idx = 0, b;   // type 0: D(b,k)   1: P(b,k,i=idx)   2: R(b,k,j=idx)
    if (ticket < B) {
      type = 0; k = 0; b = ticket;
    } else {
      const int t2 = ticket - B;
      const int s = t2 / per_step, rem = t2 - s * per_step;
      // slot q of the step's list and sample b.  The first 1 + dpos slots (the tile below the diagonal, the look-ahead
      // diagonal task and what lies between) run sample-fastest, so that the diagonal task of a sample sits B tickets
      // behind the panel tile it waits for; the rest of the step runs SLOT-fastest, sample by sample: the tiles of one
      // sample's step are then in flight together and share their common operand (block row k of L) through L2 instead
      // of fetching it from HBM once per tile (with the sample index fastest everywhere, B - 1 other samples passed
      // through L2 between two tiles of the same block row: 47 GB of DRAM reads per launch at c3, L2 hit rate 50 %).
      int q;
      {
        const int dpos_ = (s < nb - 1) ? ((1 + fa.dgap < nb - 1) ? 1 + fa.dgap : nb - 1) : -1;
        const int head = (dpos_ + 1) * B;           // tickets of the sample-fastest head of the list
        if (FUSED || rem < head) {
          // (chain launches stay sample-fastest throughout: with a handful of samples the steps of all of them should
          // advance together, and everything fits L2 anyway -- factor 1.79 ms against 1.90 ms at c2 / B = 8)
          q = rem / B;
          b = rem - q * B;
        } else {
          // (the last step has nb - 1 slots: the inverse tiles of block row nb-1)
          const int rest = (s < nb - 1) ? nb - (dpos_ + 1) : nb - 1, r2 = rem - head;
          b = r2 / rest;
          q = dpos_ + 1 + (r2 - b * rest);
        }
      }
      k = s;
      if (s < nb - 1) {
        const int dpos = (1 + fa.dgap < nb - 1) ? 1 + fa.dgap : nb - 1;
        if (q == dpos) {
          type = 0; k = s + 1;
        } else {
          const int r = q < dpos ? q : q - 1;
          if (r < nb - 1 - s) { type = 1; idx = s + 1 + r; }
          else { type = 2; idx = r - (nb - 1 - s); }
        }
      } else {
        type = 2; idx = q;
      }
    }
    const int k0 = k * TILE;
    double* L = fa.L + (int64_t)b * npad * npad;
    double* T = fa.T + (int64_t)b * npad * npad;
    int32_t* lflag = fa.lflag + (int64_t)b * nb;
    int32_t* tflag = fa.tflag + (int64_t)b * nb;
    const int rows_k = min(TILE, fa.n - k0);   // valid rows of block row k
    if (type == 0 && FUSED && k > 0) {
      // ---------------- Dpre(b,k): fused-panel launches ----------------
      // the part of a diagonal task that does not depend on block column k-1: A_kk - sum_{m<k-1} L[k,m] L[k,m]^T, left in
      // place of A_kk for the sample's chain CTA (below), which applies the last block column itself
      double* Akk = L + (int64_t)k0 * npad + k0;
      FacKK g;
      load_neg_tile(g.acc, Akk, npad, wm, wn, gq, t);
      if (k > 1) {
        const bool idle_quadrant = (wm == 0 && wn == 1);
        SlabWaiter w{lflag + k, lflag + k, fa.ctl, &s_known, 0, 0, fa.max_spins, nullptr, 0, FAC_SPB};
        g.run(smem, L + (int64_t)k0 * npad, npad, rows_k, L + (int64_t)k0 * npad, npad, 64, k0 - TILE,
              [&](int kt) { w(kt); }, idle_quadrant ? 0x7fffffff : 0);
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
          *reinterpret_cast<double2*>(Akk + (int64_t)r * npad + c) = make_double2(-g.acc[i][j][0], -g.acc[i][j][1]);
        }
      publish(fa.dflag + (int64_t)b * nb + k, 1);
      FPROF(2);
    } else if (type == 0) {
      // ---------------- D(b,k) ----------------
      // Throughput launches: one diagonal task per ticket.  Fused-panel launches (few samples: fewer tile tasks per step
      // than resident CTAs, the chain of diagonal tasks IS the run time): the CTA that draws D(b,0) becomes the CHAIN of
      // sample b and factorises every diagonal block k = 0 .. nb-1 itself, keeping T_kk in shared memory for the next
      // link: it forms the panel tile L[k,k-1] = S T_mm^T (m = k-1) from the S that P(b,m,k) accumulated and parked in
      // the still unused tile (k,m) of the T slab (sflag; R(b,k,m) writes that tile only after T_kk is published) and from
      // its own T_mm, applies it to the partial block left by Dpre(b,k) (dflag) and factorises -- between two diagonal
      // factorisations there is no flag hop, no release and no trip through L2 left, only the two products.  P(b,m,k)
      // still finishes the tile for everybody else; products and update are the same code on the same operands as in
      // throughput launches, so both modes give the same bits.  Such launches carry a third 64 x 64 tile of shared memory.
      constexpr bool chain = FUSED;
      double* blk = chain ? smem + FAC_THIRD_OFF : sA;   // the diagonal block: A_kk updated, then L_kk
      double* tin = chain ? sA : sB;                          // receives T_kk
      const int k_end = chain ? nb : k + 1;
      for (int kc = k; kc < k_end; kc++) {
        const int c0 = kc * TILE;
        const int rows_c = min(TILE, fa.n - c0);
        double* Akk = L + (int64_t)c0 * npad + c0;
        double* Tkk = T + (int64_t)c0 * npad + c0;
        FEVENT(0, kc, 0);
        if (!chain) {
          FacKK g;
          load_neg_tile(g.acc, Akk, npad, wm, wn, gq, t);
          if (kc > 1) {
            // only the lower triangle of the block is used: the warp of the upper-right quadrant computes nothing;
            // block columns 0 .. k-2 of row k through the pipeline: final long before this task is on the critical path
            const bool idle_quadrant = (wm == 0 && wn == 1);
            SlabWaiter w{lflag + kc, lflag + kc, fa.ctl, &s_known, 0, 0, fa.max_spins, nullptr, 0, FAC_SPB};
            g.run(smem, L + (int64_t)c0 * npad, npad, rows_c, L + (int64_t)c0 * npad, npad, 64, c0 - TILE,
                  [&](int kt) { w(kt); }, idle_quadrant ? 0x7fffffff : 0);
          }
          // the block as updated so far (A - sum over block columns 0 .. k-2) goes to shared memory ...
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
              *reinterpret_cast<double2*>(&blk[r * FAC_LDS + c]) = make_double2(-g.acc[i][j][0], -g.acc[i][j][1]);
            }
          if (kc > 0) {
            // ... and block column k-1, what the whole factorisation waits for (published by the panel task of the
            // previous step), is applied there: the 64 x 64 tile L[k,k-1] is fetched in ONE round of loads (sixteen
            // 16-byte loads in flight per thread, instead of four pipeline slabs issued two at a time)
            FEVENT(0, kc, 4);
            wait_flag(lflag + kc, kc, fa.ctl, fa.max_spins);
            FEVENT(0, kc, 5);
            stage_tile(sB, L + (int64_t)c0 * npad + (c0 - TILE), npad);
            __syncthreads();
            FEVENT(0, kc, 6);
          }
        } else if (kc == 0) {
          stage_tile(blk, Akk, npad);
          __syncthreads();
        } else {
          // chain link: S of the tile (k,k-1) is in sB and the partial block in blk (both parked long ago by P(b,m,k) and
          // Dpre(b,k), fetched by cp.async behind the previous link's factorisation); T_mm is still in `tin`
          FEVENT(0, kc, 4);
          cp_async_wait<0>();
          __syncthreads();
          FEVENT(0, kc, 5);
          double xa[8][2][2];
          panel_product(xa, sB, tin, warp, gq, t);
          __syncthreads();   // everyone has read S and T_mm
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int jf = c ? 7 - warp : warp;
#pragma unroll
            for (int ii = 0; ii < 8; ii++)
              *reinterpret_cast<double2*>(&sB[(8 * ii + gq) * FAC_LDS + 8 * jf + 2 * t]) = make_double2(xa[ii][c][0], xa[ii][c][1]);
          }
          __syncthreads();
          FEVENT(0, kc, 6);
        }
        if (kc > 0) {
          // the update with block column k-1 runs on the 36 8 x 8 fragments of the lower triangle only, nine per warp (the
          // 32 x 32 quadrant layout of the pipeline computes 48 fragments on three warps)
          diag_update_blocked(blk, sB, warp, gq, t);
          FPROF(2);
        }
        FEVENT(0, kc, 1);
        if (tid == 0) s_bad = __ldcg(fa.info + b);
        __syncthreads();
        FPROF(7);
        diag_chol_inv_blocked(blk, tin, s_dval, s_inv, &s_bad, c0, fa.prof);
        FPROF(6);
        FEVENT(0, kc, 2);
        // T_kk is what the panel and inverse tasks of this step wait for: stored and published first; L_kk itself is
        // read by no task of this kernel and follows behind the flags (chain: in front, so that its buffer is free)
        for (int e = tid; e < TILE * TILE / 2; e += FAC_THREADS) {
          const int r = e >> 5, c = (e & 31) * 2;
          *reinterpret_cast<double2*>(Tkk + (int64_t)r * npad + c) = *reinterpret_cast<const double2*>(&tin[r * FAC_LDS + c]);
        }
        if (tid == 0) fa.info[b] = s_bad;
        if (chain) {
          // T_kk out first (tflag[k] = k + 1 is what its consumers wait for; nothing in this kernel reads L_kk): the tasks it
          // feeds -- P(b,k,k+2), then Dpre(b,k+2) and the parked S of the link after next -- form a second chain of ~13 us
          // per step with little slack against this one
          publish(tflag + kc, kc + 1);
          for (int e = tid; e < TILE * TILE / 2; e += FAC_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            *reinterpret_cast<double2*>(Akk + (int64_t)r * npad + c) = *reinterpret_cast<const double2*>(&blk[r * FAC_LDS + c]);
          }
          // Two flags in ONE wait (their loads in flight together, one barrier): S of the next tile parked, the next partial
          // block parked -- both happened during the factorisation.  The barrier also says that nobody reads L_kk or the
          // old X any more, so the next link's operands can travel now.
          if (kc + 1 < nb) {
            SlabWaiter w3{fa.sflag + (int64_t)b * nb + kc + 1, fa.dflag + (int64_t)b * nb + kc + 1, fa.ctl, &s_known, 0, 0,
                          fa.max_spins};
            w3(0);
            stage_tile_async(sB, T + (int64_t)(c0 + TILE) * npad + c0, npad);
            stage_tile_async(blk, L + (int64_t)(c0 + TILE) * npad + (c0 + TILE), npad);
            cp_async_commit();
          }
        } else {
          publish2(lflag + kc, kc + 1, tflag + kc, kc + 1);
        }
        FEVENT(0, kc, 3);
        if (!chain) {
          for (int e = tid; e < TILE * TILE / 2; e += FAC_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            *reinterpret_cast<double2*>(Akk + (int64_t)r * npad + c) = *reinterpret_cast<const double2*>(&blk[r * FAC_LDS + c]);
          }
        }
        // sum of log L_ii and the diagonal-block share of beta = T z (row r of T_kk times z_k), both in a fixed order
        if (warp == 0) {
          double v = log(s_dval[lane]) + log(s_dval[lane + 32]);
          v = warp_sum(v);
          if (lane == 0) fa.fpart[((int64_t)b * nb + kc) * 2 + 1] = v;
        } else if (tid >= 64) {
          // z_k through s_inv (free after the factorisation; only these two warps touch it): the row sums run on
          // shared-memory operands four ahead of the FMA chain instead of one global load per term
          const int r = tid - 64;
          s_inv[r] = __ldg(fa.z + (int64_t)b * npad + c0 + r);
          asm volatile("bar.sync 1, 64;");
          const double* trow = tin + r * FAC_LDS;
          double acc = 0.0;
          int c = 0;
          for (; c + 3 <= r; c += 4) {
            const double t0 = trow[c], t1 = trow[c + 1], t2 = trow[c + 2], t3 = trow[c + 3];
            const double z0 = s_inv[c], z1 = s_inv[c + 1], z2 = s_inv[c + 2], z3 = s_inv[c + 3];
            acc = fma(t0, z0, acc);
            acc = fma(t1, z1, acc);
            acc = fma(t2, z2, acc);
            acc = fma(t3, z3, acc);
          }
          for (; c <= r; c++) acc = fma(trow[c], s_inv[c], acc);
          fa.beta[(int64_t)b * npad + c0 + r] = acc;
        }
        FPROF(5);
      }
    } else if (type == 1) {
      // ---------------- P(b,k,i) ----------------
      const int i = idx, i0 = i * TILE;
      double* Aik = L + (int64_t)i0 * npad + k0;
      if (i == k + 1) FEVENT(1, k, 0);
      // Throughput launches: a panel / inverse task of step k usually starts long after T_kk was published.  Thread 0 looks
      // at the flag when the task starts; if T_kk is there, its tile is fetched by cp.async into the pipeline stage that
      // falls free while the last slab is multiplied (one 32-deep stage holds a staged 64 x 64 tile), and S goes to the
      // other stage -- the flag's round trip and the staging round (~1.7 us in front of every epilogue) are gone.
      constexpr bool PREFETCH_TKK = !FUSED && FacKK::STAGES == 2 && FacKK::STAGE_DOUBLES >= TILE * FAC_LDS &&
                                    FacKR::STAGE_DOUBLES >= TILE * FAC_LDS;
      double* sS = sA;    // where S and T_kk are staged for the epilogue
      double* sTk = sB;
      bool have_tkk = false;
      FacKK g;
      load_neg_tile(g.acc, Aik, npad, wm, wn, gq, t);
      if (k > 0) {
        if (PREFETCH_TKK && tid == 0) s_tkk = ld_acquire(tflag + k) >= k + 1;
        SlabWaiter w{lflag + i, lflag + k, fa.ctl, &s_known, 0, 0, fa.max_spins, nullptr, 0, FAC_SPB};
        if constexpr (PREFETCH_TKK) {
          const double* Tkk_g = T + (int64_t)k0 * npad + k0;
          g.run(smem, L + (int64_t)i0 * npad, npad, min(TILE, fa.n - i0), L + (int64_t)k0 * npad, npad, 64, k0,
                [&](int kt) { w(kt); }, 0, [&](double* buf) {
                  if (s_tkk) stage_tile_async(buf, Tkk_g, npad);
                });
          have_tkk = s_tkk != 0;
          if (have_tkk) {
            const int KT = k0 / FacKK::BK;
            sTk = smem + (KT % 2) * FacKK::STAGE_DOUBLES;
            sS = smem + ((KT + 1) % 2) * FacKK::STAGE_DOUBLES;
          }
        } else {
          g.run(smem, L + (int64_t)i0 * npad, npad, min(TILE, fa.n - i0), L + (int64_t)k0 * npad, npad, 64, k0,
                [&](int kt) { w(kt); });
        }
        FPROF(2);
      }
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int r = wm * 32 + ii * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
          *reinterpret_cast<double2*>(&sS[r * FAC_LDS + c]) = make_double2(-g.acc[ii][j][0], -g.acc[ii][j][1]);
        }
      if (FUSED && i == k + 1) {
        // fused-panel launches: S goes to the unused tile (i,k) of the T slab for D(b,i) (see there)
        double* Sik = T + (int64_t)i0 * npad + k0;
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int r = wm * 32 + ii * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
            *reinterpret_cast<double2*>(Sik + (int64_t)r * npad + c) = make_double2(-g.acc[ii][j][0], -g.acc[ii][j][1]);
          }
        publish(fa.sflag + (int64_t)b * nb + i, 1);
      }
      FPROF(4);
      if (i == k + 1) FEVENT(1, k, 1);
      if (have_tkk) {
        __syncthreads();   // S is staged (T_kk arrived inside run())
      } else {
        wait_flag(tflag + k, k + 1, fa.ctl, fa.max_spins);   // T[k,k] is there (also orders the S writes)
        FPROF(3);
        if (i == k + 1) FEVENT(1, k, 2);
        stage_tile(sTk, T + (int64_t)k0 * npad + k0, npad);
        __syncthreads();
      }
      // X = S T_kk^T (panel_product), straight from the accumulators to the tile's place in L
      {
        double xa[8][2][2];
        panel_product(xa, sS, sTk, warp, gq, t);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int jf = c ? 7 - warp : warp;
#pragma unroll
          for (int ii = 0; ii < 8; ii++)
            *reinterpret_cast<double2*>(Aik + (int64_t)(8 * ii + gq) * npad + 8 * jf + 2 * t) =
                make_double2(xa[ii][c][0], xa[ii][c][1]);
        }
      }
      if (fa.fault == 1 && b == 0 && k == 0 && i == 1) continue;   // injected fault: this tile is never published
      publish(lflag + i, k + 1);
      if (i == k + 1) FEVENT(1, k, 3);
      FPROF(4);
    } else {
      // ---------------- R(b,k,j) ----------------
      if (!fa.want_inverse) continue;
      const int j = idx, j0 = j * TILE;
      // (T_kk prefetched into the free pipeline stage as in the panel tasks; here T_kk is the A operand, S the B operand)
      constexpr bool PREFETCH_TKK = !FUSED && FacKR::STAGES == 2 && FacKR::STAGE_DOUBLES >= TILE * FAC_LDS;
      double* sS = sB;
      double* sTk = sA;
      bool have_tkk = false;
      FacKR g;
      g.zero();
      if (PREFETCH_TKK && tid == 0) s_tkk = ld_acquire(tflag + k) >= k + 1;
      SlabWaiter w{lflag + k, tflag + j, fa.ctl, &s_known, j, 0, fa.max_spins, nullptr, 0, FAC_SPB};
      // first slab: B = T[j,j] is lower triangular, its columns n >= 32 vanish for the first 32 k
      if constexpr (PREFETCH_TKK) {
        const double* Tkk_g = T + (int64_t)k0 * npad + k0;
        g.run(smem, L + (int64_t)k0 * npad + j0, npad, rows_k, T + (int64_t)j0 * npad + j0, npad, 64, k0 - j0,
              [&](int kt) { w(kt); }, wn == 1 ? 32 / FAC_BK : 0, [&](double* buf) {
                if (s_tkk) stage_tile_async(buf, Tkk_g, npad);
              });
        have_tkk = s_tkk != 0;
        if (have_tkk) {
          const int KT = (k0 - j0) / FacKR::BK;
          sTk = smem + (KT % 2) * FacKR::STAGE_DOUBLES;
          sS = smem + ((KT + 1) % 2) * FacKR::STAGE_DOUBLES;
        }
      } else {
        g.run(smem, L + (int64_t)k0 * npad + j0, npad, rows_k, T + (int64_t)j0 * npad + j0, npad, 64, k0 - j0,
              [&](int kt) { w(kt); }, wn == 1 ? 32 / FAC_BK : 0);
      }
      FPROF(2);
#pragma unroll
      for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int r = wm * 32 + ii * 8 + gq, c = wn * 32 + jj * 8 + 2 * t;
          *reinterpret_cast<double2*>(&sS[r * FAC_LDS + c]) = make_double2(g.acc[ii][jj][0], g.acc[ii][jj][1]);
        }
      FPROF(4);
      if (have_tkk) {
        __syncthreads();
      } else {
        wait_flag(tflag + k, k + 1, fa.ctl, fa.max_spins);
        FPROF(3);
        stage_tile(sTk, T + (int64_t)k0 * npad + k0, npad);
        __syncthreads();
      }
      // T[k,j] = -T_kk S on 8 x 8 fragments.  T_kk is lower triangular: the fragment row fr needs k < 8 fr + 8 only.  Warp w
      // takes the fragment rows w and 7 - w (balanced as in the panel tasks) and all eight column fragments.
      {
        double xr[2][8][2];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) xr[0][jj][0] = xr[0][jj][1] = xr[1][jj][0] = xr[1][jj][1] = 0.0;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int fr = c ? 7 - warp : warp;
          const int kmax = 8 * fr + 8;
#pragma unroll 2
          for (int kk = 0; kk < kmax; kk += 4) {
            const double av = sTk[(8 * fr + gq) * FAC_LDS + kk + t];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) dmma884(xr[c][jj][0], xr[c][jj][1], av, sS[(kk + t) * FAC_LDS + 8 * jj + gq]);
          }
        }
        double* Tkj = T + (int64_t)k0 * npad + j0;
        // beta = T z: this tile's share T[k,j] z_j while the tile is in registers (one pass over T saved).  A warp holds
        // whole rows of the tile, so one 64-vector per tile goes to row j0 of the UNUSED upper tile (j, k) of the L slab,
        // from where beta_reduce_kernel adds the shares up in a fixed order.
        const double* zj = fa.z + (int64_t)b * npad + j0 + 2 * t;
        double zc[8][2];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          zc[jj][0] = __ldg(zj + 8 * jj);
          zc[jj][1] = __ldg(zj + 8 * jj + 1);
        }
        double* part = L + (int64_t)j0 * npad + k0;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int fr = c ? 7 - warp : warp;
          double sum = 0.0;
#pragma unroll
          for (int jj = 0; jj < 8; jj++) {
            *reinterpret_cast<double2*>(Tkj + (int64_t)(8 * fr + gq) * npad + 8 * jj + 2 * t) =
                make_double2(-xr[c][jj][0], -xr[c][jj][1]);
            sum = fma(xr[c][jj][0], zc[jj][0], fma(xr[c][jj][1], zc[jj][1], sum));
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          if (t == 0) part[8 * fr + gq] = -sum;
        }
      }
      publish(tflag + j, k + 1);
      FPROF(4);
    }
  }
  FPROF_FLUSH(fa.prof);
}

// beta = T z  (= L^-1 z), block row k per CTA, and the block's share of beta^T beta.  grid (nb, B), 256 threads.
__global__ void __launch_bounds__(256) beta_kernel(const double* __restrict__ Tall, const double* __restrict__ zall,
                                                   int npad, double* __restrict__ beta_all, double* __restrict__ fpart) {
  __shared__ double red[TILE];
  const int b = blockIdx.y, k = blockIdx.x, tid = threadIdx.x, nb = gridDim.x;
  const int row = tid >> 2, part = tid & 3;
  const double* Tr = Tall + (int64_t)b * npad * npad + (int64_t)(k * TILE + row) * npad;
  const double* z = zall + (int64_t)b * npad;
  double s = 0.0;
  for (int c = part * 4; c < (k + 1) * TILE; c += 16) {
    const double2 a0 = *reinterpret_cast<const double2*>(Tr + c);
    const double2 a1 = *reinterpret_cast<const double2*>(Tr + c + 2);
    const double2 z0 = *reinterpret_cast<const double2*>(z + c);
    const double2 z1 = *reinterpret_cast<const double2*>(z + c + 2);
    s += (a0.x * z0.x + a0.y * z0.y) + (a1.x * z1.x + a1.y * z1.y);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (part == 0) {
    beta_all[(int64_t)b * npad + k * TILE + row] = s;
    red[row] = s * s;
  }
  __syncthreads();
  if (tid < 32) {
    double q = red[tid] + red[tid + 32];
    q = warp_sum(q);
    if (tid == 0) fpart[((int64_t)b * nb + k) * 2] = q;
  }
}

// beta = T z from the shares the factor kernel left behind: block row k = diagonal share (in beta) + the shares of the
// tiles (k, j < k), one 64-vector each, parked in row j0 of the upper tile (j, k) of the L slab; and the
// block's share of beta^T beta.  grid (nb, B), 64 threads.  Fixed order: deterministic, independent of the schedule.
__global__ void __launch_bounds__(64) beta_reduce_kernel(const double* __restrict__ Lall, int npad,
                                                        double* __restrict__ beta_all, double* __restrict__ fpart) {
  __shared__ double red[2];
  const int b = blockIdx.y, k = blockIdx.x, r = threadIdx.x, nb = gridDim.x;
  const double* Lb = Lall + (int64_t)b * npad * npad + k * TILE + r;
  double s0 = 0.0, s1 = 0.0;   // two interleaved sums over the tiles of the block row (fixed order)
  for (int j = 0; j + 1 < k; j += 2) {
    s0 += Lb[(int64_t)(j * TILE) * npad];
    s1 += Lb[(int64_t)((j + 1) * TILE) * npad];
  }
  if (k & 1) s0 += Lb[(int64_t)((k - 1) * TILE) * npad];
  double* beta = beta_all + (int64_t)b * npad + k * TILE + r;
  const double v = (s0 + s1) + *beta;
  *beta = v;
  double q = warp_sum(v * v);
  if ((r & 31) == 0) red[r >> 5] = q;
  __syncthreads();
  if (r == 0) fpart[((int64_t)b * nb + k) * 2] = red[0] + red[1];
}

}  // namespace avn

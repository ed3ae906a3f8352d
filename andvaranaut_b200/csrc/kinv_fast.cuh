// K3 (single-kernel models): K^-1 tiles by DMMA fused with the gradient contraction, with the
// per-element work of the epilogue reduced to one kernel evaluation.
//
//   mainloop   Kinv[i,j] = sum_{k>=i} T[k,i]^T T[k,j]                       (DMMA, as the generic kernel)
//   epilogue   U   = Xs_i Xs_j^T                         gram form of r2     (DMMA, depth d)
//              WK  = (alpha_i alpha_j - Kinv) * kv * k'(r2)                 (elementwise, in registers)
//              P   = WK   [X_j | 1],   Q = WK^T [X_i | 1]                   (DMMA, 64x64x(d+1))
//              sum_ij WK_ij (x_im - x_jm)^2 = sum_i x_im^2 R_i + sum_j x_jm^2 C_j - 2 sum_i x_im P_im
//              d ll / d xw_i = 2/l^2 (x_i R_i - P_i),  d ll / d xw_j = 2/l^2 (x_j C_j - Q_j)
//   with R = P[:, d] (row sums) and C = Q[:, d] (column sums).  All O(d N^2) contractions therefore run on
//   the tensor pipe; the FP64 pipe only evaluates sqrt/exp once per matrix element.
#pragma once
#include "avn_dev.cuh"
#include "tile_gemm.cuh"

namespace avn {

struct KinvFastLayout {
  int dpad, lds, np, lda, ldp;
  int off_w, off_xsi, off_xsj, off_xai, off_xaj, off_vec, total;  // in doubles
  __host__ __device__ explicit KinvFastLayout(int d) {
    dpad = (d + 3) & ~3;
    lds = (dpad % 8 == 0) ? dpad + 4 : dpad + 8;   // ld % 16 in {4, 12}: conflict-free fragment loads
    np = (d + 1 + 7) & ~7;
    lda = np + 4;
    ldp = np + 1;
    off_w = 0;
    off_xsi = off_w + TILE * (TILE + SPAD);
    off_xsj = off_xsi + TILE * lds;
    off_xai = off_xsj + TILE * lds;
    off_xaj = off_xai + TILE * lda;
    off_vec = off_xaj + TILE * lda;
    total = off_vec + 4 * TILE;
  }
};

#ifndef AVN_KINV_BK
#define AVN_KINV_BK 16
#define AVN_KINV_STAGES 4
#endif
using KinvG2 = TileGemm<64, 64, AVN_KINV_BK, 32, 32, AVN_KINV_STAGES, true, true>;

// one tile (ti, tj) of sample b: everything described above.  smem: the CTA's dynamic shared memory; hyp / wpart: its
// static scratch.  Called once per CTA by the grid kernel and in a loop by the single-sample kernel.
template <int KIND, bool WITH_GX>
__device__ __forceinline__ void kinv_fast_tile(const KernDesc& kd, int N, int npad, const double* __restrict__ theta,
                                               const double* __restrict__ Tall, const double* __restrict__ alpha_all,
                                               const double* __restrict__ xw_all, const double* __restrict__ xs_all,
                                               const double* __restrict__ x2_all, double* __restrict__ gpart,
                                               double* __restrict__ gxpart, int b, int tile, int ntiles_all, double* smem,
                                               HypS& hyp, double (&wpart)[4][MAXACC]) {
  using G = KinvG2;
  constexpr int LDW = TILE + SPAD;
  const int tid = threadIdx.x;
  int ti, tj;
  tri_index(tile, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d;
  const double* T = Tall + (int64_t)b * npad * npad;
  load_hyp(hyp, kd, theta + (int64_t)b * kd.P);
  G g;
  g.zero();
  // the wm = 1 warps skip half of the first slab: alternate which hardware warps (SM sub-partitions) those are
  const int swz = (int)(((unsigned)tile * 2654435761u) >> 16) & 1;   // a function of the tile only: batch-size independent results
  g.swz = swz;
  // rows >= N of T are those of the identity: the k range stops at N rounded up to the slab depth, and the
  // padding rows of the last block row issue no DMMA
  // first slab: A = T[i,i] is lower triangular, its columns m >= 32 vanish for the first 32 k
  g.run(smem, T + (int64_t)i0 * npad + i0, npad, min(TILE, N - i0), T + (int64_t)i0 * npad + j0, npad, 64,
        min(npad, (N + G::BK - 1) / G::BK * G::BK) - i0, [](int) {}, (((threadIdx.x >> 5) ^ swz) % G::WARPS_M) == 1 ? 32 / G::BK : 0);

  // ---- stage the small operands (the pipeline buffers are free after run()) ----
  const KinvFastLayout lay(d);
  double* sW = smem + lay.off_w;
  double* sXsi = smem + lay.off_xsi;
  double* sXsj = smem + lay.off_xsj;
  double* sXai = smem + lay.off_xai;
  double* sXaj = smem + lay.off_xaj;
  double* sx2i = smem + lay.off_vec;
  double* sx2j = sx2i + TILE;
  double* sai = sx2j + TILE;
  double* saj = sai + TILE;
  const double* xw = xw_all + (int64_t)b * npad * d;
  const double* xs = xs_all + (int64_t)b * npad * d;   // nkern == 1
  const double* x2 = x2_all + (int64_t)b * npad;
  for (int e = tid; e < TILE * lay.dpad; e += G::NTHREADS) {
    int r = e / lay.dpad, m = e % lay.dpad;
    sXsi[r * lay.lds + m] = (m < d) ? xs[(int64_t)(i0 + r) * d + m] : 0.0;
    sXsj[r * lay.lds + m] = (m < d) ? xs[(int64_t)(j0 + r) * d + m] : 0.0;
  }
  for (int e = tid; e < TILE * lay.np; e += G::NTHREADS) {
    int r = e / lay.np, m = e % lay.np;
    sXai[r * lay.lda + m] = (m < d) ? xw[(int64_t)(i0 + r) * d + m] : (m == d ? 1.0 : 0.0);
    sXaj[r * lay.lda + m] = (m < d) ? xw[(int64_t)(j0 + r) * d + m] : (m == d ? 1.0 : 0.0);
  }
  if (tid < TILE) {
    sx2i[tid] = x2[i0 + tid];
    sx2j[tid] = x2[j0 + tid];
    sai[tid] = alpha_all[(int64_t)b * npad + i0 + tid];
    saj[tid] = alpha_all[(int64_t)b * npad + j0 + tid];
  }
  for (int e = tid; e < 4 * MAXACC; e += G::NTHREADS) (&wpart[0][0])[e] = 0.0;
  __syncthreads();

  const int warp = (tid >> 5) ^ swz, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const double symw = (ti == tj) ? 1.0 : 2.0;

  // ---- U = Xs_i Xs_j^T by DMMA ----
  double U[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) U[i][j][0] = U[i][j][1] = 0.0;
  for (int kk = 0; kk < lay.dpad; kk += 4) {
    double a[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = sXsi[(wm * 32 + i * 8 + gq) * lay.lds + kk + t];
#pragma unroll
    for (int j = 0; j < 4; j++) bb[j] = sXsj[(wn * 32 + j * 8 + gq) * lay.lds + kk + t];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma884(U[i][j][0], U[i][j][1], a[i], bb[j]);
  }

  // ---- elementwise: W, kernel value / derivative, WK tile to shared memory ----
  double trw = 0.0, skv = 0.0, sal = 0.0;
  const double kvk = hyp.kv[0], alpha = hyp.alpha;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double wkv[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t + h;
        const int I = i0 + r, J = j0 + c;
        double w = (I < N && J < N) ? (sai[r] * saj[c] - g.acc[i][j][h]) : 0.0;
        if (I == J) trw += w;
        w *= symw;
        double r2 = (sx2i[r] + sx2j[c]) - 2.0 * U[i][j][h];
        r2 = r2 > 0.0 ? r2 : 0.0;
        double kk_, dk_;
        kern_val_fast<KIND, true>(r2, alpha, kk_, dk_);
        skv = fma(w, kk_, skv);
        wkv[h] = w * kvk * dk_;
        // Exponential: k' ~ 1/r; the diagonal carries no lengthscale / input gradient ((x_i - x_j) = 0) and would
        // only amplify the rounding of the gram-form r2_ii
        if constexpr (KIND == AVN_EXPONENTIAL) wkv[h] = (I == J) ? 0.0 : wkv[h];
        if constexpr (KIND == AVN_RATQUAD) {
          double base = 1.0 + 0.5 * r2 / alpha;
          sal += w * kvk * kk_ * (-log(base) + (0.5 * r2 / alpha) / base);
        }
      }
      const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
      *reinterpret_cast<double2*>(&sW[r * LDW + c]) = make_double2(wkv[0], wkv[1]);
    }
  {
    const int slot_kv = d, slot_gv = d + 1, slot_alpha = d + 2;
    double s = warp_sum(trw);
    if (lane == 0) wpart[warp][slot_gv] = s;
    s = warp_sum(skv);
    if (lane == 0) wpart[warp][slot_kv] = s;
    if constexpr (KIND == AVN_RATQUAD) {
      s = warp_sum(sal);
      if (lane == 0) wpart[warp][slot_alpha] = s;
    }
  }
  __syncthreads();

  // ---- P = WK [X_j | 1],  Q = WK^T [X_i | 1]: warp w owns rows 16w .. 16w+15 ----
  const int NI = lay.np >> 3;
  double Pa[2][3][2], Qa[2][3][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Pa[i][j][0] = Pa[i][j][1] = Qa[i][j][0] = Qa[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < TILE; kk += 4) {
    double ap[2], aq[2], bj[3], bi[3];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int m = warp * 16 + i * 8 + gq;
      ap[i] = sW[m * LDW + kk + t];        // WK[m][k]
      aq[i] = sW[(kk + t) * LDW + m];      // WK[k][m]
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      bj[j] = (j < NI) ? sXaj[(kk + t) * lay.lda + j * 8 + gq] : 0.0;
      bi[j] = (j < NI) ? sXai[(kk + t) * lay.lda + j * 8 + gq] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        if (j < NI) {
          dmma884(Pa[i][j][0], Pa[i][j][1], ap[i], bj[j]);
          dmma884(Qa[i][j][0], Qa[i][j][1], aq[i], bi[j]);
        }
  }
  __syncthreads();  // everyone is done reading sW: reuse it for P and Q
  double* sP = sW;
  double* sQ = sW + TILE * lay.ldp;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (j < NI) {
        const int r = warp * 16 + i * 8 + gq, c = j * 8 + 2 * t;
        sP[r * lay.ldp + c] = Pa[i][j][0];
        sP[r * lay.ldp + c + 1] = Pa[i][j][1];
        sQ[r * lay.ldp + c] = Qa[i][j][0];
        sQ[r * lay.ldp + c + 1] = Qa[i][j][1];
      }
  __syncthreads();

  // ---- lengthscale slots and d ll / d xw ----
  {
    const bool rowside = tid < TILE;
    const int r = rowside ? tid : tid - TILE;
    const double* sX = rowside ? sXai : sXaj;
    const double* sS = rowside ? sP : sQ;
    const double RC = sS[r * lay.ldp + d];   // row sum (rowside) or column sum
    const int nb = npad / TILE;
    double* gx = nullptr;
    if (WITH_GX) {
      // rows of block i take this tile's row part (source tj); rows of block j its column part (source ti)
      if (rowside) gx = gxpart + (((int64_t)b * nb + tj) * npad + i0 + r) * d;
      else if (ti != tj) gx = gxpart + (((int64_t)b * nb + ti) * npad + j0 + r) * d;
    }
    for (int m = 0; m < d; m++) {
      const double il = hyp.invl[0][m], il2 = il * il;
      const double x = sX[r * lay.lda + m], pm = sS[r * lay.ldp + m];
      double v = rowside ? (x * x * RC - 2.0 * x * pm) : (x * x * RC);
      v = warp_sum(v * il2);
      if (lane == 0) wpart[warp][m] = v;
      if (WITH_GX && gx) gx[m] = 2.0 * il2 * (x * RC - pm) / symw;
    }
  }
  __syncthreads();
  const int nacc = d + 3;
  const int64_t ntiles = ntiles_all;
  double* gp = gpart + ((int64_t)b * ntiles + tile) * MAXACC;
  for (int e = tid; e < nacc; e += G::NTHREADS) gp[e] = (wpart[0][e] + wpart[1][e]) + (wpart[2][e] + wpart[3][e]);
}

// grid (ntiles_lower, B): one tile per CTA, tiles numbered longest first
template <int KIND, bool WITH_GX>
__global__ void __launch_bounds__(KinvG2::NTHREADS, 3) kinv_grad_fast_kernel(
    KernDesc kd, int N, int npad, const double* __restrict__ theta, const double* __restrict__ Tall,
    const double* __restrict__ alpha_all, const double* __restrict__ xw_all, const double* __restrict__ xs_all,
    const double* __restrict__ x2_all, double* __restrict__ gpart, double* __restrict__ gxpart) {
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  __shared__ double wpart[4][MAXACC];
  kinv_fast_tile<KIND, WITH_GX>(kd, N, npad, theta, Tall, alpha_all, xw_all, xs_all, x2_all, gpart, gxpart, blockIdx.y,
                                blockIdx.x, gridDim.x, smem, hyp, wpart);
}

// ONE sample (the latency path): a little more than one wave of tiles, whose lengths run from nb slabs down to 1.  The
// hardware deals consecutive CTAs to the SMs in a pattern of its own (blocks 0, 70 and 134 share an SM on this part: 70
// slabs on that SM against 40 on average -- the kernel took 174 us where the work is worth 110), so here a CTA takes its
// first tile by WHERE it runs: slot s = 0, 1, 2 of SM m (a counter per SM) gets tile m of round 0 walked forwards and of
// the later rounds walked backwards, which gives every SM about the same number of slabs (44 .. 41 at nb = 32); whatever
// is left (the shortest tiles) is handed out through one counter as CTAs finish.  A claim word per tile (atomicCAS)
// makes every tile run exactly once whatever the placement does.  Results are indexed by tile: same bits.
// sched: [0] tail counter, [1 .. 256] per-SM slot counters, [257] claimed tiles of the first wave, [258 ...] claim words of
// the first wave -- zeroed by the host.
template <int KIND, bool WITH_GX>
__global__ void __launch_bounds__(KinvG2::NTHREADS, 3) kinv_grad_fast_single_kernel(
    KernDesc kd, int N, int npad, const double* __restrict__ theta, const double* __restrict__ Tall,
    const double* __restrict__ alpha_all, const double* __restrict__ xw_all, const double* __restrict__ xs_all,
    const double* __restrict__ x2_all, double* __restrict__ gpart, double* __restrict__ gxpart, int ntiles, int nsm,
    int32_t* __restrict__ sched) {
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  __shared__ double wpart[4][MAXACC];
  __shared__ int s_tile;
  int32_t* claimed = sched + 258;
  const int wave = min(ntiles, (int)gridDim.x);   // tiles dealt by placement; the rest through the tail counter
  bool first = true;
  for (;;) {
    if (threadIdx.x == 0) {
      int tile = -1;
      if (first) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const int bin = (int)(smid % (unsigned)nsm);
        const int slot = atomicAdd(sched + 1 + (bin & 255), 1);
        const int base = slot * nsm, cnt = min(nsm, wave - base);
        if (bin < cnt) {
          const int t = slot == 0 ? bin : base + cnt - 1 - bin;
          if (atomicCAS(claimed + t, 0, 1) == 0) {
            tile = t;
            atomicAdd(sched + 257, 1);
          }
        }
      }
      if (tile < 0) {
        const int c = wave + atomicAdd(sched, 1);
        if (c < ntiles) tile = c;
      }
      if (tile < 0 && *reinterpret_cast<volatile int32_t*>(sched + 257) < wave) {
        // insurance: a tile of the first wave that nobody claimed (an SM numbering or placement other than the expected
        // one, or a CTA that has not started yet -- it will find its word taken and go to the tail counter)
        for (int t = 0; t < wave && tile < 0; t++)
          if (*reinterpret_cast<volatile int32_t*>(claimed + t) == 0 && atomicCAS(claimed + t, 0, 1) == 0) {
            tile = t;
            atomicAdd(sched + 257, 1);
          }
      }
      s_tile = tile;
    }
    first = false;
    __syncthreads();
    const int tile = s_tile;
    if (tile < 0) break;
    kinv_fast_tile<KIND, WITH_GX>(kd, N, npad, theta, Tall, alpha_all, xw_all, xs_all, x2_all, gpart, gxpart, 0, tile, ntiles,
                                  smem, hyp, wpart);
    __syncthreads();
  }
}

}  // namespace avn

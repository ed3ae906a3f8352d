// K3 (two-kernel folds 'A+B' / 'A*B', at most one RatQuad): the DMMA epilogue of kinv_fast.cuh with the
// product rule of the fold (reference: the kernel string parser gpmcmc.py:496-515 and the left-to-right
// fold gpmcmc.py:282-307).
//
//   K = v0 (+|*) v1,  v_q = kv_q k_q(r2_q),  c_q = dK/dv_q  (1 for '+', the other value for '*')
//   mainloop   Kinv[i,j] = sum_{k>=i} T[k,i]^T T[k,j]                               (DMMA)
//   pass q=0   U = Xs0_i Xs0_j^T (DMMA) -> k_0 and kv_0 k_0' parked in two shared tiles A, B
//   pass q=1   U = Xs1_i Xs1_j^T (DMMA) -> k_1, k_1'; then in place
//              A <- WK_0 = W c_0 kv_0 k_0',   B <- WK_1 = W c_1 kv_1 k_1'
//   per q      P = WK_q [X_j | 1],  Q = WK_q^T [X_i | 1]  (DMMA)  -> lengthscale slots of kernel q, d ll / d xw
// The Xs operands (needed for U only) and the [X | 1] operands (needed for P, Q only) share one region, so that
// two CTAs fit on an SM for every d <= 16.  Both folds commute, so a RatQuad kernel is always taken in the second pass
// (qb), where its alpha term has r2 and the fold coefficient at hand.
#pragma once
#include "avn_dev.cuh"
#include "kinv_fast.cuh"
#include "tile_gemm.cuh"

namespace avn {

struct KinvFoldLayout {
  int dpad, lds, np, lda, ldp;
  int off_a, off_b, off_x, off_vec, total;  // in doubles
  __host__ __device__ explicit KinvFoldLayout(int d) {
    dpad = (d + 3) & ~3;
    lds = (dpad % 8 == 0) ? dpad + 4 : dpad + 8;
    np = (d + 1 + 7) & ~7;
    lda = np + 4;
    ldp = np + 1;
    off_a = 0;
    off_b = off_a + TILE * (TILE + SPAD);
    off_x = off_b + TILE * (TILE + SPAD);
    const int xs = 2 * TILE * lds, xa = 2 * TILE * lda;
    off_vec = off_x + (xs > xa ? xs : xa);
    total = off_vec + 4 * TILE;
  }
};

// RQ: one of the two kernels is RatQuad (its own instantiation keeps the alpha term and the pass swap out of the
// common one)
template <bool WITH_GX, bool RQ>
__global__ void __launch_bounds__(KinvG2::NTHREADS, 2) kinv_grad_fold2_kernel(
    KernDesc kd, int N, int npad, const double* __restrict__ theta, const double* __restrict__ Tall,
    const double* __restrict__ alpha_all, const double* __restrict__ xw_all, const double* __restrict__ xs_all,
    const double* __restrict__ x2_all, double* __restrict__ gpart, double* __restrict__ gxpart) {
  using G = KinvG2;
  constexpr int LDW = TILE + SPAD;
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  __shared__ double wpart[4][MAXACC];
  const int b = blockIdx.y, tid = threadIdx.x;
  int ti, tj;
  tri_index(blockIdx.x, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d;
  const double* T = Tall + (int64_t)b * npad * npad;
  load_hyp(hyp, kd, theta + (int64_t)b * kd.P);
  G g;
  g.zero();
  const int swz = (int)((blockIdx.x * 2654435761u) >> 16) & 1;
  g.swz = swz;
  g.run(smem, T + (int64_t)i0 * npad + i0, npad, min(TILE, N - i0), T + (int64_t)i0 * npad + j0, npad, 64,
        min(npad, (N + G::BK - 1) / G::BK * G::BK) - i0, [](int) {}, (((threadIdx.x >> 5) ^ swz) % G::WARPS_M) == 1 ? 32 / G::BK : 0);
  __syncthreads();

  const KinvFoldLayout lay(d);
  double* sA = smem + lay.off_a;
  double* sB = smem + lay.off_b;
  double* sXsi = smem + lay.off_x;
  double* sXsj = sXsi + TILE * lay.lds;
  double* sXai = smem + lay.off_x;
  double* sXaj = sXai + TILE * lay.lda;
  double* sx2i = smem + lay.off_vec;
  double* sx2j = sx2i + TILE;
  double* sai = sx2j + TILE;
  double* saj = sai + TILE;
  const double* xw = xw_all + (int64_t)b * npad * d;

  const int warp = (tid >> 5) ^ swz, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const double symw = (ti == tj) ? 1.0 : 2.0;
  const bool mul = kd.op[0] != AVN_ADD;
  const int qa = (RQ && kd.kern[0] == AVN_RATQUAD) ? 1 : 0, qb = 1 - qa;   // pass order
  const double kv0 = hyp.kv[qa], kv1 = hyp.kv[qb], alpha = hyp.alpha;
  double trw = 0.0, skv0 = 0.0, skv1 = 0.0, sal = 0.0;

  for (int e = tid; e < 4 * MAXACC; e += G::NTHREADS) (&wpart[0][0])[e] = 0.0;
  if (tid < TILE) {
    sai[tid] = alpha_all[(int64_t)b * npad + i0 + tid];
    saj[tid] = alpha_all[(int64_t)b * npad + j0 + tid];
  }

  for (int p = 0; p < 2; p++) {
    const int q = p ? qb : qa;
    const double* xs = xs_all + ((int64_t)b * 2 + q) * npad * d;
    const double* x2 = x2_all + ((int64_t)b * 2 + q) * npad;
    if (p) __syncthreads();   // everyone is done with the operands of the first pass
    for (int e = tid; e < TILE * lay.dpad; e += G::NTHREADS) {
      int r = e / lay.dpad, m = e % lay.dpad;
      sXsi[r * lay.lds + m] = (m < d) ? xs[(int64_t)(i0 + r) * d + m] : 0.0;
      sXsj[r * lay.lds + m] = (m < d) ? xs[(int64_t)(j0 + r) * d + m] : 0.0;
    }
    if (tid < TILE) {
      sx2i[tid] = x2[i0 + tid];
      sx2j[tid] = x2[j0 + tid];
    }
    __syncthreads();

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

    const int kind = kd.kern[q];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
        double kq[2], dq[2], r2h[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double r2 = (sx2i[r] + sx2j[c + h]) - 2.0 * U[i][j][h];
          r2h[h] = r2 > 0.0 ? r2 : 0.0;
          kern_val(kind, r2h[h], alpha, kq[h], dq[h]);
        }
        double2* pa = reinterpret_cast<double2*>(&sA[r * LDW + c]);
        double2* pb = reinterpret_cast<double2*>(&sB[r * LDW + c]);
        if (p == 0) {
          *pa = make_double2(kq[0], kq[1]);
          *pb = make_double2(kv0 * dq[0], kv0 * dq[1]);
        } else {
          const double2 k0 = *pa, d0 = *pb;   // this thread's own elements of pass 0
          const double k0h[2] = {k0.x, k0.y}, d0h[2] = {d0.x, d0.y};
          double wk0[2], wk1[2];
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int I = i0 + r, J = j0 + c + h;
            double w = (I < N && J < N) ? (sai[r] * saj[c + h] - g.acc[i][j][h]) : 0.0;
            if (I == J) trw += w;
            w *= symw;
            const double c0 = mul ? kv1 * kq[h] : 1.0, c1 = mul ? kv0 * k0h[h] : 1.0;
            const double w0 = w * c0, w1 = w * c1;
            skv0 = fma(w0, k0h[h], skv0);
            skv1 = fma(w1, kq[h], skv1);
            if constexpr (RQ) {
              const double base = 1.0 + 0.5 * r2h[h] / alpha;
              sal += w1 * kv1 * kq[h] * (-log(base) + (0.5 * r2h[h] / alpha) / base);
            }
            // (x_i - x_j) = 0 on the diagonal: those elements carry no lengthscale / input gradient, and dropping them
            // keeps the Exponential kernel's k' ~ 1/r from amplifying the rounding of the gram-form r2_ii
            wk0[h] = (I == J) ? 0.0 : w0 * d0h[h];
            wk1[h] = (I == J) ? 0.0 : w1 * kv1 * dq[h];
          }
          *pa = make_double2(wk0[0], wk0[1]);
          *pb = make_double2(wk1[0], wk1[1]);
        }
      }
  }
  {
    const int slot_kv = 2 * d, slot_gv = 2 * d + 2, slot_alpha = 2 * d + 3;
    double s = warp_sum(trw);
    if (lane == 0) wpart[warp][slot_gv] = s;
    s = warp_sum(skv0);
    if (lane == 0) wpart[warp][slot_kv + qa] = s;
    s = warp_sum(skv1);
    if (lane == 0) wpart[warp][slot_kv + qb] = s;
    if constexpr (RQ) {
      s = warp_sum(sal);
      if (lane == 0) wpart[warp][slot_alpha] = s;
    }
  }
  __syncthreads();   // WK tiles complete; the Xs operands are dead
  for (int e = tid; e < TILE * lay.np; e += G::NTHREADS) {
    int r = e / lay.np, m = e % lay.np;
    sXai[r * lay.lda + m] = (m < d) ? xw[(int64_t)(i0 + r) * d + m] : (m == d ? 1.0 : 0.0);
    sXaj[r * lay.lda + m] = (m < d) ? xw[(int64_t)(j0 + r) * d + m] : (m == d ? 1.0 : 0.0);
  }
  __syncthreads();

  // ---- per kernel: P = WK [X_j | 1],  Q = WK^T [X_i | 1]; warp w owns rows 16w .. 16w+15 ----
  const int NI = lay.np >> 3;
  for (int q = 0; q < 2; q++) {
    double* sW = q ? sB : sA;
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
        ap[i] = sW[m * LDW + kk + t];
        aq[i] = sW[(kk + t) * LDW + m];
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
    __syncthreads();  // everyone is done reading this WK tile: its storage takes P and Q of the kernel
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
  }
  __syncthreads();

  // ---- lengthscale slots of both kernels and d ll / d xw ----
  {
    const bool rowside = tid < TILE;
    const int r = rowside ? tid : tid - TILE;
    const double* sX = rowside ? sXai : sXaj;
    const double* s0 = rowside ? sA : sA + TILE * lay.ldp;
    const double* s1 = rowside ? sB : sB + TILE * lay.ldp;
    const double RC0 = s0[r * lay.ldp + d], RC1 = s1[r * lay.ldp + d];
    const int nb = npad / TILE;
    double* gx = nullptr;
    if (WITH_GX) {
      if (rowside) gx = gxpart + (((int64_t)b * nb + tj) * npad + i0 + r) * d;
      else if (ti != tj) gx = gxpart + (((int64_t)b * nb + ti) * npad + j0 + r) * d;
    }
    for (int m = 0; m < d; m++) {
      const double il0 = hyp.invl[qa][m], il1 = hyp.invl[qb][m];
      const double a0 = il0 * il0, a1 = il1 * il1;
      const double x = sX[r * lay.lda + m], p0 = s0[r * lay.ldp + m], p1 = s1[r * lay.ldp + m];
      double v0 = rowside ? (x * x * RC0 - 2.0 * x * p0) : (x * x * RC0);
      double v1 = rowside ? (x * x * RC1 - 2.0 * x * p1) : (x * x * RC1);
      v0 = warp_sum(v0 * a0);
      v1 = warp_sum(v1 * a1);
      if (lane == 0) {
        wpart[warp][qa * d + m] = v0;
        wpart[warp][qb * d + m] = v1;
      }
      if (WITH_GX && gx) gx[m] = (2.0 * a0 * (x * RC0 - p0) + 2.0 * a1 * (x * RC1 - p1)) / symw;
    }
  }
  __syncthreads();
  const int nacc = 2 * d + 4;
  const int64_t ntiles = gridDim.x;
  double* gp = gpart + ((int64_t)b * ntiles + blockIdx.x) * MAXACC;
  for (int e = tid; e < nacc; e += G::NTHREADS) gp[e] = (wpart[0][e] + wpart[1][e]) + (wpart[2][e] + wpart[3][e]);
}

}  // namespace avn

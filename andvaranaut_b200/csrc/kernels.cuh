// Kernels of the GP inner loop (sm_100a).  Layout conventions:
//   * every N x N matrix of a hyperparameter sample b is stored row-major in a padded
//     [npad][npad] slab (npad = N rounded up to 64), lower block-triangle valid; padding rows/cols
//     carry the identity so that L, T = L^-1 stay block-diag(., I) and no kernel needs ragged tiles;
//   * the batch dimension is the slowest one; blockIdx.y (or .x for per-sample kernels) = b.
#pragma once
#include "avn_dev.cuh"
#include "tile_gemm.cuh"
#include "warp.cuh"
#include "factor.cuh"

namespace avn {

struct WsPtrs {
  double* xw;      // [B][npad][d]            warped inputs
  double* dxw;     // [B][npad][d][MAXWP]     d warped input / d warp params (only with learnable x warps)
  double* xs;      // [B][nkern][npad][d]     inputs scaled by 1/l of each kernel
  double* x2;      // [B][nkern][npad]        row norms of xs (NumPy summation order)
  double* z;       // [B][npad]               converted outputs
  double* dz;      // [B][npad][MAXWP]        d z / d output-warp params
  double* wstat;   // [B][16]                 0: sum log g'; 1..8: its param derivatives
  double* kl;      // [B][npad][npad]         K then L
  double* t;       // [B][npad][npad]         T = L^-1
  double* beta;    // [B][npad]
  double* alpha;   // [B][npad]
  double* gpart;   // [B][ntiles][MAXACC]     per-tile partial sums of the gradient contraction
  double* gxpart;  // [B][nb][npad][d]        per-source-tile partial sums of d ll / d warped input
  double* fpart;   // [B][nb][2]              per block row: beta_k^T beta_k, sum log diag(L_kk)
  int32_t* lflag;  // [B][nb]                 progress flags of the factor kernel (factor.cuh)
  int32_t* tflag;  // [B][nb]
  int32_t* sflag;  // [B][nb]                 behind ctl
  int32_t* dflag;  // [B][nb]                 behind sflag
  int32_t* ksched; // scheduler words of kinv_grad_fast_single_kernel (behind dflag; zeroed with the flags)
  int32_t* ctl;    // [8]                     ticket counter, abort flag
};

constexpr int WSTAT = 16;

// ------------------------------------------------------------------------------------------------
// K0: conversions.  grid (d + 1, B): CTA m < d converts input column m, CTA d converts the outputs.
// ------------------------------------------------------------------------------------------------
constexpr int WARP_THREADS = 256;   // (512 threads: 85 -> 74 us at B = 1 but slower batched, and the block reductions change order)
__global__ void __launch_bounds__(WARP_THREADS) warp_kernel(KernDesc kd, WarpProgs progs, const double* __restrict__ X,
                                                   const double* __restrict__ y, int N, int npad,
                                                   const double* __restrict__ theta, WsPtrs ws, int stage_doubles,
                                                   int col_begin) {
  // stage_doubles >= N (1 + nparams): the column and its parameter Jacobians live in dynamic shared memory while the
  // stages of the warp program run over them (every stage is a pass over the column, the data-dependent ones several
  // with block reductions between: on the strided global arrays each pass paid an L2 round trip per element -- 90 us of
  // a 1.1 ms evaluation at B = 1) and are written to their global layout once at the end.  Same arithmetic either way.
  extern __shared__ __align__(16) double wsm[];
  __shared__ double sh[192];
  // col_begin: first column of this launch (0: all of them; d: only the output column, launched beside the input columns)
  const int b = blockIdx.y, m = blockIdx.x + col_begin, tid = threadIdx.x, nt = blockDim.x;
  const double* th = theta + (int64_t)b * kd.P;
  if (m < kd.d) {
    double* xw = ws.xw + (int64_t)b * npad * kd.d + m;
    const avn_warp_prog& pr = progs.xw[m];
    const bool staged = pr.nstages > 0 && stage_doubles >= N * (1 + pr.nparams);
    for (int n = tid; n < npad; n += nt) {
      const double v = (n < N) ? X[(int64_t)n * kd.d + m] : 0.0;
      if (staged && n < N) wsm[n] = v;
      else xw[(int64_t)n * kd.d] = v;
    }
    if (pr.nstages > 0) {
      int poff = 0;
      for (int q = 0; q < m; q++)
        if (progs.xw[q].nstages > 0) poff += progs.xw[q].nparams;
      __syncthreads();
      double dummy = 0, ddummy[MAXWP];
      double* dual = ws.dxw + ((int64_t)b * npad * kd.d + m) * MAXWP;
      if (staged) {
        const int np = pr.nparams;
        double* sdual = wsm + N;
        run_warp_column(pr, th + kd.off_iw + poff, N, wsm, 1, sdual, np, 0, dummy, ddummy, sh);
        __syncthreads();
        for (int n = tid; n < N; n += nt) {
          xw[(int64_t)n * kd.d] = wsm[n];
          for (int q = 0; q < np; q++) dual[(int64_t)n * kd.d * MAXWP + q] = sdual[n * np + q];
        }
      } else {
        run_warp_column(pr, th + kd.off_iw + poff, N, xw, kd.d, dual, (int64_t)kd.d * MAXWP, 0, dummy, ddummy, sh);
      }
    }
    return;
  }
  double* z = ws.z + (int64_t)b * npad;
  const bool staged = progs.yw.nstages > 0 && stage_doubles >= N * (1 + progs.yw.nparams);
  for (int n = tid; n < npad; n += nt) {
    const double v = (n < N) ? y[n] : 0.0;
    if (staged && n < N) wsm[n] = v;
    else z[n] = v;
  }
  __syncthreads();
  double* wst = ws.wstat + (int64_t)b * WSTAT;
  if (progs.yw.nstages > 0) {
    double lsum = 0, dlsum[MAXWP];
    for (int q = 0; q < MAXWP; q++) dlsum[q] = 0.0;
    double* dz = ws.dz + (int64_t)b * npad * MAXWP;
    if (staged) {
      const int np = progs.yw.nparams;
      double* sdual = wsm + N;
      run_warp_column(progs.yw, th + kd.off_cw, N, wsm, 1, sdual, np, 1, lsum, dlsum, sh);
      __syncthreads();
      for (int n = tid; n < N; n += nt) {
        z[n] = wsm[n];
        for (int q = 0; q < np; q++) dz[(int64_t)n * MAXWP + q] = sdual[n * np + q];
      }
    } else {
      run_warp_column(progs.yw, th + kd.off_cw, N, z, 1, dz, MAXWP, 1, lsum, dlsum, sh);
    }
    double tot = block_sum(lsum, sh);
    if (tid == 0) wst[0] = tot;
    for (int q = 0; q < progs.yw.nparams; q++) {
      double tq = block_sum(dlsum[q], sh);
      if (tid == 0) wst[1 + q] = tq;
    }
  } else if (tid == 0) {
    wst[0] = 0.0;
  }
}

// scaled copies and row norms per kernel (X * (1/ls); sum(square(.), 1)).  grid (npad / 256, B), one row per thread.
__global__ void __launch_bounds__(256) scale_kernel(KernDesc kd, int npad, const double* __restrict__ theta, WsPtrs ws) {
  __shared__ double invl[MAXK][MAXD];
  const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
  const double* th = theta + (int64_t)b * kd.P;
  for (int i = threadIdx.x; i < kd.nkern * kd.d; i += 256) invl[i / kd.d][i % kd.d] = 1.0 / th[kd.off_l + i];
  __syncthreads();
  if (n >= npad) return;
  const double* xw = ws.xw + ((int64_t)b * npad + n) * kd.d;
  double xr[MAXD];
  for (int mm = 0; mm < kd.d; mm++) xr[mm] = xw[mm];
  for (int k = 0; k < kd.nkern; k++) {
    double* xs = ws.xs + (((int64_t)b * kd.nkern + k) * npad + n) * kd.d;
    double tmp[MAXD];
    for (int mm = 0; mm < kd.d; mm++) {
      tmp[mm] = __dmul_rn(xr[mm], invl[k][mm]);
      xs[mm] = tmp[mm];
    }
    ws.x2[((int64_t)b * kd.nkern + k) * npad + n] = sumsq_numpy_order(tmp, kd.d);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: covariance build, lower 64x64 tiles, batched.  grid (ntiles_lower, B), 256 threads.
// Writes K + (gv + jitter) I; padding rows/cols carry the identity.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tri_index(int t, int& i, int& j) {
  i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((i + 1) * (i + 2) / 2 <= t) i++;
  while (i * (i + 1) / 2 > t) i--;
  j = t - i * (i + 1) / 2;
}

#ifndef AVN_COV_MINB
#define AVN_COV_MINB 4   // 64 registers: measured 8 % faster than 2 CTAs of 128 registers (latency-bound exp chains)
#endif
__global__ void __launch_bounds__(256, AVN_COV_MINB) cov_kernel(KernDesc kd, int N, int npad, const double* __restrict__ theta,
                                                  const double* __restrict__ xs_all, const double* __restrict__ x2_all,
                                                  double* __restrict__ Kout) {
  extern __shared__ __align__(16) double smem[];
  __shared__ HypS hyp;
  const int b = blockIdx.y, tid = threadIdx.x;
  int ti, tj;
  tri_index(blockIdx.x, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d, nk = kd.nkern;
  load_hyp(hyp, kd, theta + (int64_t)b * kd.P);
  // smem, dimension-major so that a thread's 4 rows / 4 columns are one 32-byte read:
  //   xi[nk][d][64], xj[nk][d][64], x2i[nk][64], x2j[nk][64]
  double* sxi = smem;
  double* sxj = sxi + nk * d * TILE;
  double* s2i = sxj + nk * d * TILE;
  double* s2j = s2i + nk * TILE;
  for (int k = 0; k < nk; k++) {
    const double* xs = xs_all + ((int64_t)b * nk + k) * npad * d;
    const double* x2 = x2_all + ((int64_t)b * nk + k) * npad;
    for (int e = tid; e < TILE * d; e += 256) {
      const int r = e / d, m = e % d;
      sxi[(k * d + m) * TILE + r] = xs[(int64_t)i0 * d + e];
      sxj[(k * d + m) * TILE + r] = xs[(int64_t)j0 * d + e];
    }
    if (tid < TILE) {
      s2i[k * TILE + tid] = x2[i0 + tid];
      s2j[k * TILE + tid] = x2[j0 + tid];
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double* Kb = Kout + (int64_t)b * npad * npad;
  const double dadd = hyp.gv + kd.jitter;
  double out[4][4];
  for (int k = 0; k < nk; k++) {
    double dot[4][4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) dot[rr][cc] = 0.0;
    for (int m = 0; m < d; m++) {
      const double2* pi = reinterpret_cast<const double2*>(sxi + (k * d + m) * TILE + ty * 4);
      const double2* pj = reinterpret_cast<const double2*>(sxj + (k * d + m) * TILE + tx * 4);
      const double2 a0 = pi[0], a1 = pi[1], b0 = pj[0], b1 = pj[1];
      const double xi[4] = {a0.x, a0.y, a1.x, a1.y}, xj[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int rr = 0; rr < 4; rr++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) dot[rr][cc] = fma(xi[rr], xj[cc], dot[rr][cc]);   // sequential in m
    }
    const int kind = kd.kern[k];
    const double kvk = hyp.kv[k];
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        double r2 = __dadd_rn(__dmul_rn(-2.0, dot[rr][cc]), __dadd_rn(s2i[k * TILE + ty * 4 + rr], s2j[k * TILE + tx * 4 + cc]));
        r2 = r2 > 0.0 ? r2 : 0.0;
        const double v = __dmul_rn(kvk, kern_val_only(kind, r2, hyp.alpha));
        if (k == 0) out[rr][cc] = v;
        else out[rr][cc] = (kd.op[k - 1] == AVN_ADD) ? __dadd_rn(out[rr][cc], v) : __dmul_rn(out[rr][cc], v);
      }
  }
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int I = i0 + ty * 4 + rr;
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      const int J = j0 + tx * 4 + cc;
      if (I >= N || J >= N) out[rr][cc] = (I == J) ? 1.0 : 0.0;
      else if (I == J) out[rr][cc] += dadd;
    }
    double2* dst = reinterpret_cast<double2*>(Kb + (int64_t)I * npad + j0 + tx * 4);
    dst[0] = make_double2(out[rr][0], out[rr][1]);
    dst[1] = make_double2(out[rr][2], out[rr][3]);
  }
}

// Single-kernel models (nkern == 1), kernel kind known at compile time: the same tile and thread mapping, but the
// element loop holds no switch, no kernel fold and the trimmed exp (avn_dev.cuh).
template <int KIND>
__global__ void __launch_bounds__(256, AVN_COV_MINB) cov1_kernel(KernDesc kd, int N, int npad, const double* __restrict__ theta,
                                                                 const double* __restrict__ xs_all,
                                                                 const double* __restrict__ x2_all, double* __restrict__ Kout) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, tid = threadIdx.x;
  int ti, tj;
  tri_index(blockIdx.x, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d;
  const double* th = theta + (int64_t)b * kd.P;
  const double kvk = th[kd.off_kv];
  const double dadd = (kd.noise ? th[kd.off_gv] : 0.0) + kd.jitter;
  const double alpha = kd.has_alpha ? th[kd.off_alpha] : 1.0;
  double* sxi = smem;                 // [d][64]
  double* sxj = sxi + d * TILE;       // [d][64]
  double* s2i = sxj + d * TILE;
  double* s2j = s2i + TILE;
  const double* xs = xs_all + (int64_t)b * npad * d;
  const double* x2 = x2_all + (int64_t)b * npad;
  for (int e = tid; e < TILE * d; e += 256) {
    const int r = e / d, m = e - r * d;
    sxi[m * TILE + r] = xs[(int64_t)i0 * d + e];
    sxj[m * TILE + r] = xs[(int64_t)j0 * d + e];
  }
  if (tid < TILE) s2i[tid] = x2[i0 + tid];
  else if (tid < 2 * TILE) s2j[tid - TILE] = x2[j0 + tid - TILE];
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double dot[4][4];
#pragma unroll
  for (int rr = 0; rr < 4; rr++)
#pragma unroll
    for (int cc = 0; cc < 4; cc++) dot[rr][cc] = 0.0;
  for (int m = 0; m < d; m++) {
    const double2* pi = reinterpret_cast<const double2*>(sxi + m * TILE + ty * 4);
    const double2* pj = reinterpret_cast<const double2*>(sxj + m * TILE + tx * 4);
    const double2 a0 = pi[0], a1 = pi[1], b0 = pj[0], b1 = pj[1];
    const double xi[4] = {a0.x, a0.y, a1.x, a1.y}, xj[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) dot[rr][cc] = fma(xi[rr], xj[cc], dot[rr][cc]);   // sequential in m
  }
  double n2i[4], n2j[4];
  {
    const double2* pi = reinterpret_cast<const double2*>(s2i + ty * 4);
    const double2* pj = reinterpret_cast<const double2*>(s2j + tx * 4);
    const double2 a0 = pi[0], a1 = pi[1], b0 = pj[0], b1 = pj[1];
    n2i[0] = a0.x; n2i[1] = a0.y; n2i[2] = a1.x; n2i[3] = a1.y;
    n2j[0] = b0.x; n2j[1] = b0.y; n2j[2] = b1.x; n2j[3] = b1.y;
  }
  double* Kb = Kout + (int64_t)b * npad * npad;
  const bool interior = (ti != tj) && (i0 + TILE <= N);   // no diagonal, no padding (j0 < i0)
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int I = i0 + ty * 4 + rr;
    double out[4];
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      double r2 = __dadd_rn(__dmul_rn(-2.0, dot[rr][cc]), __dadd_rn(n2i[rr], n2j[cc]));
      r2 = r2 > 0.0 ? r2 : 0.0;
      double kk, dk;
      kern_val_fast<KIND, false>(r2, alpha, kk, dk);
      out[cc] = __dmul_rn(kvk, kk);
    }
    if (!interior) {
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const int J = j0 + tx * 4 + cc;
        if (I >= N || J >= N) out[cc] = (I == J) ? 1.0 : 0.0;
        else if (I == J) out[cc] += dadd;
      }
    }
    double2* dst = reinterpret_cast<double2*>(Kb + (int64_t)I * npad + j0 + tx * 4);
    dst[0] = make_double2(out[0], out[1]);
    dst[1] = make_double2(out[2], out[3]);
  }
}

// Two-kernel folds ('RBF+Matern52', 'Matern32*RatQuad', ...), both kernel kinds known at compile time: the tile and thread
// mapping of cov1_kernel, one pass per kernel (its own scaled inputs and row norms), the fold (+ or *, warp-uniform) on
// the 4 x 4 register block.  The generic cov_kernel runs a switch per element and kernel and spills at its 64-register
// budget: 1.56 ms against 2 x 0.58 ms for the two single-kernel builds at N = 2000, d = 8, B = 64.
template <int K0, int K1>
__global__ void __launch_bounds__(256, 3) cov2_kernel(KernDesc kd, int N, int npad, const double* __restrict__ theta,
                                                      const double* __restrict__ xs_all, const double* __restrict__ x2_all,
                                                      double* __restrict__ Kout) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, tid = threadIdx.x;
  int ti, tj;
  tri_index(blockIdx.x, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d;
  const double* th = theta + (int64_t)b * kd.P;
  const double kv0 = th[kd.off_kv], kv1 = th[kd.off_kv + 1];
  const double dadd = (kd.noise ? th[kd.off_gv] : 0.0) + kd.jitter;
  const double alpha = kd.has_alpha ? th[kd.off_alpha] : 1.0;
  const bool mul = kd.op[0] == AVN_MUL;
  double* sxi = smem;                     // [2][d][64]
  double* sxj = sxi + 2 * d * TILE;       // [2][d][64]
  double* s2i = sxj + 2 * d * TILE;       // [2][64]
  double* s2j = s2i + 2 * TILE;
  for (int k = 0; k < 2; k++) {
    const double* xs = xs_all + ((int64_t)b * 2 + k) * npad * d;
    const double* x2 = x2_all + ((int64_t)b * 2 + k) * npad;
    for (int e = tid; e < TILE * d; e += 256) {
      const int r = e / d, m = e - r * d;
      sxi[(k * d + m) * TILE + r] = xs[(int64_t)i0 * d + e];
      sxj[(k * d + m) * TILE + r] = xs[(int64_t)j0 * d + e];
    }
    if (tid < TILE) s2i[k * TILE + tid] = x2[i0 + tid];
    else if (tid < 2 * TILE) s2j[k * TILE + tid - TILE] = x2[j0 + tid - TILE];
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double out[4][4];
  auto pass = [&](auto kind_tag, int k, double kvk) {
    constexpr int KIND = decltype(kind_tag)::value;
    double dot[4][4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) dot[rr][cc] = 0.0;
    for (int m = 0; m < d; m++) {
      const double2* pi = reinterpret_cast<const double2*>(sxi + (k * d + m) * TILE + ty * 4);
      const double2* pj = reinterpret_cast<const double2*>(sxj + (k * d + m) * TILE + tx * 4);
      const double2 a0 = pi[0], a1 = pi[1], b0 = pj[0], b1 = pj[1];
      const double xi[4] = {a0.x, a0.y, a1.x, a1.y}, xj[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int rr = 0; rr < 4; rr++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) dot[rr][cc] = fma(xi[rr], xj[cc], dot[rr][cc]);   // sequential in m
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        double r2 = __dadd_rn(__dmul_rn(-2.0, dot[rr][cc]), __dadd_rn(s2i[k * TILE + ty * 4 + rr], s2j[k * TILE + tx * 4 + cc]));
        r2 = r2 > 0.0 ? r2 : 0.0;
        double kk, dk;
        kern_val_fast<KIND, false>(r2, alpha, kk, dk);
        const double v = __dmul_rn(kvk, kk);
        if (k == 0) out[rr][cc] = v;
        else out[rr][cc] = mul ? __dmul_rn(out[rr][cc], v) : __dadd_rn(out[rr][cc], v);
      }
  };
  pass(std::integral_constant<int, K0>{}, 0, kv0);
  pass(std::integral_constant<int, K1>{}, 1, kv1);
  double* Kb = Kout + (int64_t)b * npad * npad;
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int I = i0 + ty * 4 + rr;
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      const int J = j0 + tx * 4 + cc;
      if (I >= N || J >= N) out[rr][cc] = (I == J) ? 1.0 : 0.0;
      else if (I == J) out[rr][cc] += dadd;
    }
    double2* dst = reinterpret_cast<double2*>(Kb + (int64_t)I * npad + j0 + tx * 4);
    dst[0] = make_double2(out[rr][0], out[rr][1]);
    dst[1] = make_double2(out[rr][2], out[rr][3]);
  }
}

// K2 (Cholesky + triangular inverse + beta) lives in factor.cuh.

// alpha = T^T beta.  grid (nb, B), ALPHA_THREADS threads: 64 columns x 16 row groups (row group g takes rows j0 + g,
// j0 + g + 16, ...), eight loads in flight per thread, the sums in a fixed order (sequential per thread, then a fixed tree
// over the row groups): independent of B.  With 4 row groups a single evaluation (32 CTAs, 16 MB of T in L2) was bound
// by its own load latency at 35 us.
constexpr int ALPHA_THREADS = 1024, ALPHA_RG = ALPHA_THREADS / TILE;
__global__ void __launch_bounds__(ALPHA_THREADS) alpha_kernel(const double* __restrict__ Tall, const double* __restrict__ beta_all,
                                                              int npad, double* __restrict__ alpha_all) {
  __shared__ double part[ALPHA_RG][TILE];
  const int b = blockIdx.y, j0 = blockIdx.x * TILE, tid = threadIdx.x;
  const int c = tid & 63, rg = tid >> 6;
  const double* T = Tall + (int64_t)b * npad * npad;
  const double* beta = beta_all + (int64_t)b * npad;
  double acc = 0.0;
  int i = j0 + rg;
  for (; i + 7 * ALPHA_RG < npad; i += 8 * ALPHA_RG) {
    double tv[8], bv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      tv[u] = T[(int64_t)(i + ALPHA_RG * u) * npad + j0 + c];
      bv[u] = beta[i + ALPHA_RG * u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) acc += tv[u] * bv[u];
  }
  for (; i < npad; i += ALPHA_RG) acc += T[(int64_t)i * npad + j0 + c] * beta[i];
  part[rg][c] = acc;
  __syncthreads();
  if (tid < TILE) {
    double s8[8];
#pragma unroll
    for (int u = 0; u < 8; u++) s8[u] = part[2 * u][tid] + part[2 * u + 1][tid];
    alpha_all[(int64_t)b * npad + j0 + tid] = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// K3: K^-1 tiles (T^T T) fused with the gradient contraction  1/2 sum_ij W_ij dK_ij/dtheta,
// W = alpha alpha^T - K^-1.  K^-1 is never written to memory.
// grid (ntiles_lower, B), 128 threads; tile (i, j<=i):  Kinv[i,j] = sum_{k >= i} T[k,i]^T T[k,j].
// ------------------------------------------------------------------------------------------------
using KinvG = TileGemm<64, 64, 16, 32, 32, 4, true, true>;

template <bool WITH_GX>
__global__ void __launch_bounds__(KinvG::NTHREADS) kinv_grad_kernel(KernDesc kd, int N, int npad,
                                                                    const double* __restrict__ theta,
                                                                    const double* __restrict__ Tall,
                                                                    const double* __restrict__ alpha_all,
                                                                    const double* __restrict__ xw_all,
                                                                    double* __restrict__ gpart,
                                                                    double* __restrict__ gxpart) {
  using G = KinvG;
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  __shared__ double wpart[4][MAXACC];
  const int b = blockIdx.y, tid = threadIdx.x;
  int ti, tj;
  tri_index(blockIdx.x, ti, tj);
  const int i0 = ti * TILE, j0 = tj * TILE;
  const int d = kd.d, nk = kd.nkern;
  const double* T = Tall + (int64_t)b * npad * npad;
  load_hyp(hyp, kd, theta + (int64_t)b * kd.P);
  G g;
  g.zero();
  g.run(smem, T + (int64_t)i0 * npad + i0, npad, min(TILE, N - i0), T + (int64_t)i0 * npad + j0, npad, 64,
        min(npad, (N + G::BK - 1) / G::BK * G::BK) - i0);
  // stage x rows and alpha of both blocks (aliases the pipeline buffers, free after run())
  const int ldx = d | 1;
  double* sxi = smem;
  double* sxj = sxi + TILE * ldx;
  double* sai = sxj + TILE * ldx;
  double* saj = sai + TILE;
  double* sgr = saj + TILE;               // [2 (wn)][64][MAXD] row sums (WITH_GX), one owner lane per slot
  double* sgc = sgr + 2 * TILE * MAXD;    // [2 (wm)][64][MAXD] col sums
  const double* xw = xw_all + (int64_t)b * npad * d;
  for (int e = tid; e < TILE * d; e += G::NTHREADS) {
    int r = e / d, m = e % d;
    sxi[r * ldx + m] = xw[(int64_t)(i0 + r) * d + m];
    sxj[r * ldx + m] = xw[(int64_t)(j0 + r) * d + m];
  }
  if (tid < TILE) {
    sai[tid] = alpha_all[(int64_t)b * npad + i0 + tid];
    saj[tid] = alpha_all[(int64_t)b * npad + j0 + tid];
  }
  if (WITH_GX)
    for (int e = tid; e < 4 * TILE * MAXD; e += G::NTHREADS) sgr[e] = 0.0;
  for (int e = tid; e < 4 * MAXACC; e += G::NTHREADS) (&wpart[0][0])[e] = 0.0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const double symw = (ti == tj) ? 1.0 : 2.0;
  // W (weighted, masked) replaces the accumulators; trace part for d/d gv
  double trw = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t + h;
        int I = i0 + r, J = j0 + c;
        double w = (I < N && J < N) ? (sai[r] * saj[c] - g.acc[i][j][h]) : 0.0;
        if (I == J) trw += w;
        g.acc[i][j][h] = w * symw;
      }
  const int slot_gv = nk * d + nk, slot_alpha = slot_gv + 1;
  {
    double s = warp_sum(trw);
    if (lane == 0) wpart[warp][slot_gv] = s;
  }
  double wk[4][4][2];
  for (int k = 0; k < nk; k++) {
    double skv = 0.0, sal = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          int r = wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t + h;
          // values of all kernels at this pair (direct differences), fold coefficient of kernel k
          double vals[MAXK], kk_k = 0.0, dk_k = 0.0, r2_k = 0.0;
          for (int q = 0; q < nk; q++) {
            double r2 = 0.0;
            for (int m = 0; m < d; m++) {
              double s = (sxi[r * ldx + m] - sxj[c * ldx + m]) * hyp.invl[q][m];
              r2 = fma(s, s, r2);
            }
            double kv_, dk_;
            kern_val(kd.kern[q], r2, hyp.alpha, kv_, dk_);
            vals[q] = hyp.kv[q] * kv_;
            if (q == k) { kk_k = kv_; dk_k = dk_; r2_k = r2; }
          }
          // coef = dK/dvals[k] through the left-to-right fold
          double coef = 1.0;
          {
            double pref = vals[0];
            double prefs[MAXK];
            prefs[0] = pref;
            for (int q = 1; q < nk; q++) {
              pref = (kd.op[q - 1] == AVN_ADD) ? pref + vals[q] : pref * vals[q];
              prefs[q] = pref;
            }
            double gg = 1.0;
            for (int q = nk - 1; q >= 1; q--) {
              if (kd.op[q - 1] == AVN_ADD) {
                if (q == k) coef = gg;
              } else {
                if (q == k) coef = gg * prefs[q - 1];
                gg *= vals[q];
              }
            }
            if (k == 0) coef = gg;
          }
          double w = g.acc[i][j][h] * coef;
          skv += w * kk_k;
          wk[i][j][h] = w * hyp.kv[k] * dk_k;
          if (kd.kern[k] == AVN_RATQUAD) {
            double base = 1.0 + 0.5 * r2_k / hyp.alpha;
            double kq = pow(base, -hyp.alpha);
            sal += w * hyp.kv[k] * kq * (-log(base) + (0.5 * r2_k / hyp.alpha) / base);
          }
        }
    {
      double s = warp_sum(skv);
      if (lane == 0) wpart[warp][nk * d + k] = s;
      if (kd.kern[k] == AVN_RATQUAD) {
        double s2 = warp_sum(sal);
        if (lane == 0) wpart[warp][slot_alpha] = s2;
      }
    }
    for (int m = 0; m < d; m++) {
      const double il = hyp.invl[k][m];
      double xi[4], xj[4][2];
#pragma unroll
      for (int i = 0; i < 4; i++) xi[i] = sxi[(wm * 32 + i * 8 + gq) * ldx + m];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        xj[j][0] = sxj[(wn * 32 + j * 8 + 2 * t) * ldx + m];
        xj[j][1] = sxj[(wn * 32 + j * 8 + 2 * t + 1) * ldx + m];
      }
      double s2 = 0.0;
      double rs[4] = {0, 0, 0, 0}, cs[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            double s = (xi[i] - xj[j][h]) * il;
            double ws_ = wk[i][j][h] * s;
            s2 = fma(ws_, s, s2);
            if (WITH_GX) {
              rs[i] += ws_;
              cs[j][h] += ws_;
            }
          }
      s2 = warp_sum(s2);
      if (lane == 0) wpart[warp][k * d + m] = s2;
      if (WITH_GX) {
        // d ll / d xw[I][m] += 2 * invl * sum_J (W K')_IJ s_IJ ; the symmetric weight is removed again
        const double f = 2.0 * il / symw;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double v = rs[i];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (t == 0) sgr[(wn * TILE + wm * 32 + i * 8 + gq) * MAXD + m] += v * f;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int h = 0; h < 2; h++) {
            double v = cs[j][h];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (gq == 0) sgc[(wm * TILE + wn * 32 + j * 8 + 2 * t + h) * MAXD + m] -= v * f;
          }
      }
    }
  }
  __syncthreads();
  const int nacc = nk * d + nk + 2;
  const int64_t ntiles = gridDim.x;
  double* gp = gpart + ((int64_t)b * ntiles + blockIdx.x) * MAXACC;
  for (int e = tid; e < nacc; e += G::NTHREADS) gp[e] = (wpart[0][e] + wpart[1][e]) + (wpart[2][e] + wpart[3][e]);
  if (WITH_GX) {
    const int nb = npad / TILE;
    // rows of block i take this tile's row sums (source tj); rows of block j take its column sums (source ti)
    double* gr = gxpart + (((int64_t)b * nb + tj) * npad + i0) * d;
    for (int e = tid; e < TILE * d; e += G::NTHREADS)
      gr[e] = sgr[(e / d) * MAXD + e % d] + sgr[(TILE + e / d) * MAXD + e % d];
    if (ti != tj) {
      double* gc = gxpart + (((int64_t)b * nb + ti) * npad + j0) * d;
      for (int e = tid; e < TILE * d; e += G::NTHREADS)
        gc[e] = sgc[(e / d) * MAXD + e % d] + sgc[(TILE + e / d) * MAXD + e % d];
    }
  }
}

// d ll / d warped input: sum of the per-source-tile partials (fixed order) into slab 0.  grid (npad*d/256, B).
__global__ void __launch_bounds__(256) gx_reduce_kernel(int npad, int d, double* __restrict__ gxpart) {
  const int nb = npad / TILE;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x, slab = (int64_t)npad * d;
  if (e >= slab) return;
  double* G = gxpart + (int64_t)blockIdx.y * nb * slab;
  double gsum = 0.0;
  for (int s = 0; s < nb; s++) gsum += G[(int64_t)s * slab + e];
  G[e] = gsum;
}

// ------------------------------------------------------------------------------------------------
// finalisation: ll and gradient assembly.  grid (B, 1 + n_iw + n_cw), 256 threads: CTA y = 0 forms ll and the kernel
// hyperparameter slots, every following CTA takes ONE (dimension, parameter) pair of the learnable input warps or one
// output-warp parameter: its eight warps sum eight contiguous row ranges (eight rows in flight per lane: the strided
// reads are L2-latency bound), combined in a fixed order.  At B = 1 this kernel is pure latency: one CTA for everything
// took 61 us, one warp per pair on three CTAs 34 us.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int finalize_grid_y(int n_iw, int n_cw, int want_grad) {
  return want_grad ? 1 + n_iw + n_cw : 1;
}

__global__ void __launch_bounds__(256) finalize_kernel(KernDesc kd, WarpProgs progs, int N, int npad, int ntiles,
                                                       int want_grad, const double* __restrict__ theta, WsPtrs ws,
                                                       int32_t* __restrict__ info, double* __restrict__ ll,
                                                       double* __restrict__ grad) {
  __shared__ double slots[MAXACC];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = theta + (int64_t)b * kd.P;
  const double* wst = ws.wstat + (int64_t)b * WSTAT;
  // a wait of the factor kernel timed out (factor.cuh): nothing downstream of it can be trusted.  Reported as
  // info = -1 with ll = NaN and a zero gradient -- distinct from a non-positive pivot (info > 0, ll = -inf).
  const bool aborted = ws.ctl[1] != 0;
  const bool bad = aborted || info[b] != 0;
  const int role = blockIdx.y;   // 0: ll + kernel hyperparameters; 1 .. n_iw: input-warp pairs; then output-warp parameters
  const int warp = tid >> 5, lane = tid & 31;
  const int d = kd.d, nk = kd.nkern;
  double* gr = want_grad ? grad + (int64_t)b * kd.P : nullptr;
  if (role == 0) {
    if (aborted) {
      // info[b] is rewritten here: the other CTAs of the sample take `bad` from the abort flag itself
      if (tid == 0) {
        info[b] = -1;
        ll[b] = NAN;
      }
    } else if (tid == 0) {
      const double norm = -0.5 * N * 1.8378770664093454835606594728112;  // log(2 pi)
      const int nbk = npad / TILE;
      const double* fp = ws.fpart + (int64_t)b * nbk * 2;
      double quad = 0.0, logdet = 0.0;   // fixed order: deterministic
      for (int k = 0; k < nbk; k++) {
        quad += fp[2 * k];
        logdet += fp[2 * k + 1];
      }
      double v = (norm - 0.5 * quad) - logdet + wst[0];
      ll[b] = bad ? -INFINITY : v;
    }
    if (!want_grad) return;
    if (bad) {
      for (int e = tid; e < kd.off_iw; e += 256) gr[e] = 0.0;
      if (tid == 0 && kd.has_alpha) gr[kd.off_alpha] = 0.0;
      return;
    }
    const int nacc = nk * d + nk + 2;
    for (int e = warp; e < nacc; e += 8) {  // one warp per slot, lanes stride over tiles (fixed order: deterministic)
      double s = 0.0;
      const double* gp = ws.gpart + (int64_t)b * ntiles * MAXACC + e;
      for (int tI = lane; tI < ntiles; tI += 32) s += gp[(int64_t)tI * MAXACC];
      s = warp_sum(s);
      if (lane == 0) slots[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < nk * d; e += 256) gr[kd.off_l + e] = -slots[e] / th[kd.off_l + e];
    if (tid < nk) gr[kd.off_kv + tid] = 0.5 * slots[nk * d + tid];
    if (tid == 0) {
      if (kd.noise) gr[kd.off_gv] = 0.5 * slots[nk * d + nk];
      if (kd.has_alpha) gr[kd.off_alpha] = 0.5 * slots[nk * d + nk + 1];
    }
    return;
  }
  if (!want_grad) return;
  __shared__ double wred[8];
  if (bad) {
    // this CTA's entry of the warp-parameter gradient
    if (tid == 0) {
      if (role <= kd.n_iw) gr[kd.off_iw + role - 1] = 0.0;
      else gr[kd.off_cw + role - 1 - kd.n_iw] = 0.0;
    }
    return;
  }
  // rows of this warp: eight contiguous ranges of whole 32-row groups
  const int per = ((N + 255) / 256) * 32, n0 = warp * per, n1 = min(N, n0 + per);
  double acc;
  if (role <= kd.n_iw) {
    // learnable input warps: sum_n G[n][m] * d xw[n][m] / d p
    const int nb = npad / TILE, pq = role - 1;
    // G[n][m] = sum over source tiles, already reduced into slab 0 by gx_reduce_kernel
    const double* G = ws.gxpart + (int64_t)b * nb * npad * d;
    int m = 0, q = pq;   // pair index -> (dimension m, parameter q of its warp)
    for (;; m++) {
      const int np = progs.xw[m].nstages > 0 ? progs.xw[m].nparams : 0;
      if (q < np) break;
      q -= np;
    }
    const double* Gm = G + m;
    const double* Dm = ws.dxw + (((int64_t)b * npad) * d + m) * MAXWP + q;
    double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int n = n0 + lane;
    for (; n + 224 < n1; n += 256) {
      double gv[8], dv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        gv[u] = Gm[(int64_t)(n + 32 * u) * d];
        dv[u] = Dm[(int64_t)(n + 32 * u) * d * MAXWP];
      }
#pragma unroll
      for (int u = 0; u < 8; u++) a[u] = fma(gv[u], dv[u], a[u]);
    }
    for (; n < n1; n += 32) a[0] = fma(Gm[(int64_t)n * d], Dm[(int64_t)n * d * MAXWP], a[0]);
    acc = warp_sum(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])));
  } else {
    // learnable output warp: -alpha^T dz/dp (+ sum d log g'/dp below)
    const double* al = ws.alpha + (int64_t)b * npad;
    const double* Dz = ws.dz + (int64_t)b * npad * MAXWP + (role - 1 - kd.n_iw);
    double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int n = n0 + lane;
    for (; n + 224 < n1; n += 256) {
#pragma unroll
      for (int u = 0; u < 8; u++) a[u] = fma(al[n + 32 * u], Dz[(int64_t)(n + 32 * u) * MAXWP], a[u]);
    }
    for (; n < n1; n += 32) a[0] = fma(al[n], Dz[(int64_t)n * MAXWP], a[0]);
    acc = warp_sum(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])));
  }
  if (lane == 0) wred[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    const double tot = ((wred[0] + wred[1]) + (wred[2] + wred[3])) + ((wred[4] + wred[5]) + (wred[6] + wred[7]));
    if (role <= kd.n_iw) gr[kd.off_iw + role - 1] = tot;
    else gr[kd.off_cw + role - 1 - kd.n_iw] = -tot + wst[1 + role - 1 - kd.n_iw];
  }
}

// ------------------------------------------------------------------------------------------------
// K4: predict.  (a) cross-covariance panel K_xs [npad][mld] + latent mean  mu = K_xs^T alpha
//               (b) V = T K_xs by DMMA row tile by row tile, column sums of V^2, variance,
//                   Gauss-Hermite reversion / EI epilogue (gpmcmc.py:545-569).
// ------------------------------------------------------------------------------------------------
struct PredState {
  // layout of the state buffer produced by factorize (all offsets in doubles from the base)
  int64_t off_alpha, off_xs, off_x2, off_t, off_hyp;
};

__global__ void __launch_bounds__(256) kxs_kernel(KernDesc kd, int N, int npad, const HypS* __restrict__ hyp_g,
                                                  const double* __restrict__ xs_tr, const double* __restrict__ x2_tr,
                                                  const double* __restrict__ alpha, const double* __restrict__ Xtest,
                                                  int64_t M, int64_t m_begin, int mld, double* __restrict__ Kxs,
                                                  double* __restrict__ mu, int nsplit, double* __restrict__ mu_part) {
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  __shared__ double part[4][TILE];
  const int tid = threadIdx.x, c = tid & 63, rg = tid >> 6;
  const int d = kd.d, nk = kd.nkern;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;  // column inside the panel
  const int64_t mg = m_begin + col0 + c;            // global test index
  for (int e = tid; e < (int)(sizeof(HypS) / sizeof(double)); e += 256)
    reinterpret_cast<double*>(&hyp)[e] = reinterpret_cast<const double*>(hyp_g)[e];
  __syncthreads();
  const int ldx = d | 1;
  double* sx = smem;                 // [nk][64][ldx] scaled test points
  double* sx2 = sx + nk * TILE * ldx;  // [nk][64]
  if (rg == 0) {
    for (int k = 0; k < nk; k++) {
      double tmp[MAXD];
      for (int m = 0; m < d; m++) {
        double x = (mg < M) ? Xtest[mg * d + m] : 0.0;
        tmp[m] = __dmul_rn(x, hyp.invl[k][m]);
        sx[(k * TILE + c) * ldx + m] = tmp[m];
      }
      sx2[k * TILE + c] = sumsq_numpy_order(tmp, d);
    }
  }
  __syncthreads();
  // small test batches: the training rows are split over blockIdx.y (nsplit chunks of whole 64-row blocks)
  const int nbk = npad / TILE, per = (nbk + nsplit - 1) / nsplit;
  const int r0 = min(npad, (int)blockIdx.y * per * TILE), r1 = min(npad, r0 + per * TILE);
  double acc = 0.0;
  for (int n = r0 + rg; n < r1; n += 4) {
    double v = 0.0;
    if (n < N)
      v = cov_fold(kd, hyp, xs_tr + (int64_t)n * d, (int64_t)npad * d, x2_tr + n, npad, sx + c * ldx, TILE * ldx,
                   sx2 + c, TILE);
    Kxs[(int64_t)n * mld + col0 + c] = v;
    acc = fma(v, alpha[n], acc);
  }
  part[rg][c] = acc;
  __syncthreads();
  if (tid < TILE) {
    const double sum = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    if (nsplit > 1) mu_part[(int64_t)blockIdx.y * mld + col0 + tid] = sum;
    else if (m_begin + col0 + tid < M) mu[m_begin + col0 + tid] = sum;
  }
}

// Single-kernel models, kernel kind known at compile time: the cross-covariance panel in 4 x 4 register blocks (the
// tile and thread mapping of cov1_kernel), training rows staged block by block in shared memory with the next block
// prefetched into registers, no switch / fold in the element loop.  Same contract as kxs_kernel (grid (nblk, nsplit),
// 256 threads); the latent mean is summed per thread over its rows in rising order and then over the 16 row groups
// of the CTA in fixed order.  The one-element-at-a-time kernel above ran at 7 % of the HBM write rate (FP64 pipe
// 15 % active, latency-bound on its L1 reads); this one is bound by the FP64 pipe (sqrt + exp per element).
// shared: xt [d][64] test points, x2t [64], xr [d][64] training rows of the current block, x2r [64], al [64],
//         part [16][64]
__host__ __device__ inline size_t kxs1_smem_doubles(int d) { return (size_t)(2 * d * TILE + 3 * TILE + 16 * TILE); }

template <int KIND>
__global__ void __launch_bounds__(256, 2) kxs1_kernel(KernDesc kd, int N, int npad, const HypS* __restrict__ hyp_g,
                                                      const double* __restrict__ xs_tr, const double* __restrict__ x2_tr,
                                                      const double* __restrict__ alpha, const double* __restrict__ Xtest,
                                                      int64_t M, int64_t m_begin, int mld, double* __restrict__ Kxs,
                                                      double* __restrict__ mu, int nsplit, double* __restrict__ mu_part) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, d = kd.d;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  double* sxt = smem;                  // [d][64]
  double* s2t = sxt + d * TILE;        // [64]
  double* sxr = s2t + TILE;            // [d][64]
  double* s2r = sxr + d * TILE;        // [64]
  double* sal = s2r + TILE;            // [64]
  double* part = sal + TILE;           // [16][64]
  const double kvk = hyp_g->kv[0], kalpha = hyp_g->alpha;
  if (tid < TILE) {
    const int64_t mg = m_begin + col0 + tid;
    double tmp[MAXD];
    for (int m = 0; m < d; m++) {
      const double x = (mg < M) ? Xtest[mg * d + m] : 0.0;
      tmp[m] = __dmul_rn(x, hyp_g->invl[0][m]);
      sxt[m * TILE + tid] = tmp[m];
    }
    s2t[tid] = sumsq_numpy_order(tmp, d);
  }
  const int nbk = npad / TILE, per = (nbk + nsplit - 1) / nsplit;
  const int r0 = min(npad, (int)blockIdx.y * per * TILE), r1 = min(npad, r0 + per * TILE);
  const int tx = tid & 15, ty = tid >> 4;
  // register prefetch of one training block: 64 d scaled inputs (<= 4 per thread), 64 row norms, 64 alphas
  constexpr int PF = (MAXD * TILE + 255) / 256;
  double pf[PF], pfv = 0.0;
  auto fetch = [&](int n0) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int e = tid + 256 * u;
      pf[u] = (e < TILE * d) ? xs_tr[(int64_t)n0 * d + e] : 0.0;
    }
    if (tid < TILE) pfv = x2_tr[n0 + tid];
    else if (tid < 2 * TILE) pfv = alpha[n0 + tid - TILE];
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int e = tid + 256 * u;
      if (e < TILE * d) {
        const int r = e / d, m = e - r * d;
        sxr[m * TILE + r] = pf[u];
      }
    }
    if (tid < TILE) s2r[tid] = pfv;
    else if (tid < 2 * TILE) sal[tid - TILE] = pfv;
  };
  double macc[4] = {0.0, 0.0, 0.0, 0.0};
  if (r0 < r1) fetch(r0);
  for (int n0 = r0; n0 < r1; n0 += TILE) {
    __syncthreads();   // the previous block's shared rows are no longer read (first pass: test points staged)
    stash();
    __syncthreads();
    if (n0 + TILE < r1) fetch(n0 + TILE);
    double dot[4][4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int cc = 0; cc < 4; cc++) dot[rr][cc] = 0.0;
    for (int m = 0; m < d; m++) {
      const double2* pi = reinterpret_cast<const double2*>(sxr + m * TILE + ty * 4);
      const double2* pj = reinterpret_cast<const double2*>(sxt + m * TILE + tx * 4);
      const double2 a0 = pi[0], a1 = pi[1], b0 = pj[0], b1 = pj[1];
      const double xi[4] = {a0.x, a0.y, a1.x, a1.y}, xj[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int rr = 0; rr < 4; rr++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) dot[rr][cc] = fma(xi[rr], xj[cc], dot[rr][cc]);   // sequential in m
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
      const int n = n0 + ty * 4 + rr;
      const double n2 = s2r[ty * 4 + rr], al = sal[ty * 4 + rr];
      double out[4];
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        double r2 = __dadd_rn(__dmul_rn(-2.0, dot[rr][cc]), __dadd_rn(n2, s2t[tx * 4 + cc]));
        r2 = r2 > 0.0 ? r2 : 0.0;
        double kk, dk;
        kern_val_fast<KIND, false>(r2, kalpha, kk, dk);
        out[cc] = (n < N) ? __dmul_rn(kvk, kk) : 0.0;
        macc[cc] = fma(out[cc], al, macc[cc]);
      }
      double2* dst = reinterpret_cast<double2*>(Kxs + (int64_t)n * mld + col0 + tx * 4);
      dst[0] = make_double2(out[0], out[1]);
      dst[1] = make_double2(out[2], out[3]);
    }
  }
#pragma unroll
  for (int cc = 0; cc < 4; cc++) part[ty * TILE + tx * 4 + cc] = macc[cc];
  __syncthreads();
  if (tid < TILE) {
    double sum = 0.0;
#pragma unroll
    for (int g = 0; g < 16; g++) sum += part[g * TILE + tid];
    if (nsplit > 1) mu_part[(int64_t)blockIdx.y * mld + col0 + tid] = sum;
    else if (m_begin + col0 + tid < M) mu[m_begin + col0 + tid] = sum;
  }
}

// Gauss-Hermite reversion / EI / normvar of one point (gpmcmc.py:545-569) from the latent mean and variance
__device__ __forceinline__ void gh_epilogue(const avn_epilogue& epi, double madd, double& mu, double& var) {
  if (epi.mode == 0) return;
  const double sd = sqrt(2.0 * var);
  double s1 = 0.0, s2 = 0.0;
  for (int q = 0; q < epi.deg; q++) {
    double yi = sd * epi.nodes[q] + mu;
    double yr = prog_rev_const(epi.yrev, yi) + madd;
    double f = yr;
    if (epi.mode == 2) {
      double df = epi.ei_max ? (yr - epi.yopt) : (epi.yopt - yr);
      f = df > 0.0 ? df : 0.0;
    }
    s1 += epi.weights[q] * f;
    s2 += epi.weights[q] * (yr * yr);
  }
  const double ispi = 0.56418958354775628694807945156077;  // 1/sqrt(pi)
  mu = ispi * s1;
  var = ispi * s2 - mu * mu;
  if (epi.normvar) var /= mu * mu;
}

// row-split runs (nsplit > 1): fixed-order sum of the partial means / column sums of V^2, then the epilogue.
// grid (cols / 256), one test point per thread.
__global__ void __launch_bounds__(256) predict_finish_kernel(KernDesc kd, const HypS* __restrict__ hyp_g, int nsplit,
                                                             const double* __restrict__ mu_part,
                                                             const double* __restrict__ vpart, int mld, int64_t M,
                                                             int64_t m_begin, avn_epilogue epi, int pred_noise,
                                                             const double* __restrict__ mean_add,
                                                             double* __restrict__ mu_out, double* __restrict__ var_out) {
  const int64_t col = (int64_t)blockIdx.x * 256 + threadIdx.x, mg = m_begin + col;
  if (col >= mld || mg >= M) return;
  double mu = 0.0, vv = 0.0;
  for (int s = 0; s < nsplit; s++) {
    mu += mu_part[(int64_t)s * mld + col];
    vv += vpart[(int64_t)s * mld + col];
  }
  const HypS& hyp = *hyp_g;
  double var = kdiag_total(kd, hyp) - vv;
  if (pred_noise) var += hyp.gv;
  gh_epilogue(epi, mean_add ? mean_add[mg] : 0.0, mu, var);
  mu_out[mg] = mu;
  var_out[mg] = var;
}

#ifndef AVN_PRED_BK
#define AVN_PRED_BK 16
#define AVN_PRED_STAGES 4
#define AVN_PRED_CTAS 2      // resident CTAs per SM the panel width and the row split are sized for
#endif
using PredG = TileGemm<64, 64, AVN_PRED_BK, 32, 32, AVN_PRED_STAGES, false, true>;

__global__ void __launch_bounds__(PredG::NTHREADS) predict_var_kernel(KernDesc kd, int npad,
                                                                      const HypS* __restrict__ hyp_g,
                                                                      const double* __restrict__ T,
                                                                      const double* __restrict__ Kxs, int mld,
                                                                      int64_t M, int64_t m_begin, avn_epilogue epi,
                                                                      const double* __restrict__ mean_add,
                                                                      double* __restrict__ mu_io,
                                                                      double* __restrict__ var_out, int nsplit,
                                                                      double* __restrict__ vpart) {
  using G = PredG;
  extern __shared__ double smem[];
  __shared__ double colsq[2][TILE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  const int nb = npad / TILE;
  double cs[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  G g;
  // row blocks interleaved over blockIdx.y (cost of block ib is ib + 1 slabs); nsplit == 1: all of them
  for (int ib = blockIdx.y; ib < nb; ib += nsplit) {
    g.zero();
    g.run(smem, T + (int64_t)ib * TILE * npad, npad, 64, Kxs + col0, mld, 64, (ib + 1) * TILE);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        cs[j][0] = fma(g.acc[i][j][0], g.acc[i][j][0], cs[j][0]);
        cs[j][1] = fma(g.acc[i][j][1], g.acc[i][j][1], cs[j][1]);
      }
  }
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v = cs[j][h];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (gq == 0) colsq[wm][wn * 32 + j * 8 + 2 * t + h] = v;
    }
  __syncthreads();
  if (tid < TILE) {
    if (nsplit > 1) {
      vpart[(int64_t)blockIdx.y * mld + col0 + tid] = colsq[0][tid] + colsq[1][tid];
      return;
    }
    const int64_t mg = m_begin + col0 + tid;
    if (mg < M) {
      const HypS& hyp = *hyp_g;
      double var = kdiag_total(kd, hyp) - (colsq[0][tid] + colsq[1][tid]);
      var += hyp.gv;
      double mu = mu_io[mg];
      gh_epilogue(epi, mean_add ? mean_add[mg] : 0.0, mu, var);
      mu_io[mg] = mu;
      var_out[mg] = var;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4': predict with gradients w.r.t. the (converted) query points -- the BO refine graph of the reference
// (gpmcmc.py:738-801: kstar, v = L^-1 kstar, mean = kstar^T alpha, var = k** - v^T v, Gauss-Hermite reversion,
// EI; differentiated by PyTensor inside pm.find_MAP over x).  Here analytically:
//   d mu / d x_m = sum_i alpha_i dk_i/dx_m,   d s2 / d x_m = -2 sum_i w_i dk_i/dx_m,   w = T^T (T k*) = K^-1 k*,
//   dk_i/dx_m = sum_q coef_q kv_q k_q'(r2_q) 2 (xs*_qm - xs_qim) / l_qm   (coef_q: product rule of the kernel fold)
// and the chain rule through the reversion epilogue.
//   (a) kxs_kernel           K_xs panel + latent mean                     (as predict)
//   (b) predict_v_kernel     V = T K_xs stored, latent variance           (DMMA)
//   (c) ttv_kernel           W = T^T V                                    (DMMA), overwrites the K_xs panel
//   (d) predict_grad_kernel  the O(N d) contractions per point + epilogue chain
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PredG::NTHREADS) predict_v_kernel(KernDesc kd, int npad, const HypS* __restrict__ hyp_g,
                                                                    const double* __restrict__ T,
                                                                    const double* __restrict__ Kxs, int mld, int64_t M,
                                                                    int64_t m_begin, int pred_noise,
                                                                    double* __restrict__ V, double* __restrict__ var_out,
                                                                    int nsplit, double* __restrict__ vpart) {
  using G = PredG;
  extern __shared__ double smem[];
  __shared__ double colsq[2][TILE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  const int nb = npad / TILE;
  double cs[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  G g;
  for (int ib = blockIdx.y; ib < nb; ib += nsplit) {
    g.zero();
    g.run(smem, T + (int64_t)ib * TILE * npad, npad, 64, Kxs + col0, mld, 64, (ib + 1) * TILE);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        cs[j][0] = fma(g.acc[i][j][0], g.acc[i][j][0], cs[j][0]);
        cs[j][1] = fma(g.acc[i][j][1], g.acc[i][j][1], cs[j][1]);
        const int r = ib * TILE + wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
        *reinterpret_cast<double2*>(V + (int64_t)r * mld + col0 + c) = make_double2(g.acc[i][j][0], g.acc[i][j][1]);
      }
  }
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v = cs[j][h];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (gq == 0) colsq[wm][wn * 32 + j * 8 + 2 * t + h] = v;
    }
  __syncthreads();
  if (tid < TILE) {
    if (nsplit > 1) {
      vpart[(int64_t)blockIdx.y * mld + col0 + tid] = colsq[0][tid] + colsq[1][tid];
      return;
    }
    const int64_t mg = m_begin + col0 + tid;
    if (mg < M) {
      const HypS& hyp = *hyp_g;
      double var = kdiag_total(kd, hyp) - (colsq[0][tid] + colsq[1][tid]);
      if (pred_noise) var += hyp.gv;
      var_out[mg] = var;
    }
  }
}

using TtvG = TileGemm<64, 64, 16, 32, 32, 4, true, true>;

// W[i,:] = sum_{k >= i} T[k,i]^T V[k,:].  grid (column blocks, nb), 128 threads.
__global__ void __launch_bounds__(TtvG::NTHREADS) ttv_kernel(int npad, const double* __restrict__ T,
                                                             const double* __restrict__ V, int mld,
                                                             double* __restrict__ W) {
  using G = TtvG;
  extern __shared__ double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % G::WARPS_M, wn = warp / G::WARPS_M, gq = lane >> 2, t = lane & 3;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  const int i0 = blockIdx.y * TILE;
  G g;
  g.zero();
  g.run(smem, T + (int64_t)i0 * npad + i0, npad, 64, V + (int64_t)i0 * mld + col0, mld, 64, npad - i0);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = i0 + wm * 32 + i * 8 + gq, c = wn * 32 + j * 8 + 2 * t;
      *reinterpret_cast<double2*>(W + (int64_t)r * mld + col0 + c) = make_double2(g.acc[i][j][0], g.acc[i][j][1]);
    }
}

// chain rule of the reversion epilogue for one point: latent (mu, var) and the raw contractions
// g1[m] = sum_n alpha_n dk_n/dx_m, g2[m] = sum_n w_n dk_n/dx_m  ->  outputs and their gradients
__device__ __forceinline__ void grad_epilogue(const avn_epilogue& epi, int d, int64_t mg,
                                              const double* __restrict__ mean_add,
                                              const double* __restrict__ dmean_add, const double* g1v,
                                              const double* g2v, double* __restrict__ mu_io,
                                              double* __restrict__ var_io, double* __restrict__ dmean_out,
                                              double* __restrict__ dvar_out) {
  double mu = mu_io[mg], var = var_io[mg];
  // outputs (om, ov) as functions of the latent (mu, var): partial derivatives a* = d om, b* = d ov
  double am = 1.0, av = 0.0, bm = 0.0, bv = 1.0, cm = 0.0, cv = 0.0;   // c*: factor of d mean_add / d x
  if (epi.mode != 0) {
    const double madd = mean_add ? mean_add[mg] : 0.0;
    const double sd = sqrt(2.0 * var);
    double s1 = 0.0, s2 = 0.0, s1m = 0.0, s1v = 0.0, s2m = 0.0, s2v = 0.0, s1c = 0.0, s2c = 0.0;
    for (int q = 0; q < epi.deg; q++) {
      const double yi = sd * epi.nodes[q] + mu;
      double dyr;
      const double yr = prog_rev_const_d(epi.yrev, yi, dyr) + madd;
      double f = yr, df = 1.0;
      if (epi.mode == 2) {
        const double dff = epi.ei_max ? (yr - epi.yopt) : (epi.yopt - yr);
        f = dff > 0.0 ? dff : 0.0;
        df = dff > 0.0 ? (epi.ei_max ? 1.0 : -1.0) : 0.0;
      }
      const double w = epi.weights[q];
      const double dyv = dyr * epi.nodes[q] / sd;      // d yr / d var
      s1 += w * f;
      s2 += w * (yr * yr);
      s1m += w * df * dyr;
      s1v += w * df * dyv;
      s1c += w * df;
      s2m += w * 2.0 * yr * dyr;
      s2v += w * 2.0 * yr * dyv;
      s2c += w * 2.0 * yr;
    }
    const double ispi = 0.56418958354775628694807945156077;  // 1/sqrt(pi)
    const double om = ispi * s1;
    double ov = ispi * s2 - om * om;
    am = ispi * s1m; av = ispi * s1v; cm = ispi * s1c;
    bm = ispi * s2m - 2.0 * om * am;
    bv = ispi * s2v - 2.0 * om * av;
    cv = ispi * s2c - 2.0 * om * cm;
    if (epi.normvar) {
      const double i2 = 1.0 / (om * om), i3 = 2.0 * ov / (om * om * om);
      bm = bm * i2 - i3 * am;
      bv = bv * i2 - i3 * av;
      cv = cv * i2 - i3 * cm;
      ov *= i2;
    }
    mu = om;
    var = ov;
  }
  mu_io[mg] = mu;
  var_io[mg] = var;
  for (int m = 0; m < d; m++) {
    const double g1 = g1v[m], g2 = -2.0 * g2v[m];   // g2: d var / d x_m
    const double dm_add = dmean_add ? dmean_add[mg * d + m] : 0.0;
    dmean_out[mg * d + m] = am * g1 + av * g2 + cm * dm_add;
    dvar_out[mg * d + m] = bm * g1 + bv * g2 + cv * dm_add;
  }
}

// grid (column blocks, nsplit), 256 threads = 4 row groups x 64 test points.  nsplit > 1 (small test batches): the
// training rows are split over blockIdx.y, the CTA stores its partial contractions and predict_grad_finish_kernel
// runs the epilogue.
__global__ void __launch_bounds__(256) predict_grad_kernel(KernDesc kd, int N, int npad, const HypS* __restrict__ hyp_g,
                                                           const double* __restrict__ xs_tr,
                                                           const double* __restrict__ x2_tr,
                                                           const double* __restrict__ alpha,
                                                           const double* __restrict__ Wm, int mld,
                                                           const double* __restrict__ Xtest, int64_t M, int64_t m_begin,
                                                           avn_epilogue epi, const double* __restrict__ mean_add,
                                                           const double* __restrict__ dmean_add,
                                                           double* __restrict__ mu_io, double* __restrict__ var_io,
                                                           double* __restrict__ dmean_out, double* __restrict__ dvar_out,
                                                           int nsplit, double* __restrict__ gpart) {
  extern __shared__ double smem[];
  __shared__ HypS hyp;
  const int tid = threadIdx.x, c = tid & 63, rg = tid >> 6;
  const int d = kd.d, nk = kd.nkern;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  const int64_t mg = m_begin + col0 + c;
  for (int e = tid; e < (int)(sizeof(HypS) / sizeof(double)); e += 256)
    reinterpret_cast<double*>(&hyp)[e] = reinterpret_cast<const double*>(hyp_g)[e];
  __syncthreads();
  const int ldx = d | 1;
  double* sx = smem;                       // [nk][64][ldx] scaled test points
  double* sx2 = sx + nk * TILE * ldx;      // [nk][64]
  double* sred = sx2 + nk * TILE;          // [4][64][2 * d] partial sums of the row groups
  if (rg == 0) {
    for (int k = 0; k < nk; k++) {
      double tmp[MAXD];
      for (int m = 0; m < d; m++) {
        const double x = (mg < M) ? Xtest[mg * d + m] : 0.0;
        tmp[m] = __dmul_rn(x, hyp.invl[k][m]);
        sx[(k * TILE + c) * ldx + m] = tmp[m];
      }
      sx2[k * TILE + c] = sumsq_numpy_order(tmp, d);
    }
  }
  __syncthreads();
  double gmu[MAXD], gs[MAXD];
#pragma unroll
  for (int m = 0; m < MAXD; m++) gmu[m] = gs[m] = 0.0;
  const int nbk = npad / TILE, per = (nbk + nsplit - 1) / nsplit;
  const int r0 = min(N, (int)blockIdx.y * per * TILE), r1 = min(N, r0 + per * TILE);
  for (int n = r0 + rg; n < r1; n += 4) {
    // kernel values / derivatives of every kernel of the fold at the pair (train n, test c)
    double vals[MAXK], dks[MAXK];
    for (int q = 0; q < nk; q++) {
      const double r2 = sqdist_gram(xs_tr + ((int64_t)q * npad + n) * d, sx + (q * TILE + c) * ldx,
                                    x2_tr[(int64_t)q * npad + n], sx2[q * TILE + c], d);
      double kq, dkq;
      kern_val(kd.kern[q], r2, hyp.alpha, kq, dkq);
      vals[q] = hyp.kv[q] * kq;
      dks[q] = (r2 > 0.0) ? hyp.kv[q] * dkq : 0.0;   // the clip of square_dist has zero slope where it is active
    }
    // coef[q] = d fold / d vals[q] through the left-to-right fold (as in kinv_grad_kernel)
    double coef[MAXK];
    {
      double prefs[MAXK];
      prefs[0] = vals[0];
      for (int q = 1; q < nk; q++) prefs[q] = (kd.op[q - 1] == AVN_ADD) ? prefs[q - 1] + vals[q] : prefs[q - 1] * vals[q];
      double gg = 1.0;
      for (int q = nk - 1; q >= 1; q--) {
        if (kd.op[q - 1] == AVN_ADD) {
          coef[q] = gg;
        } else {
          coef[q] = gg * prefs[q - 1];
          gg *= vals[q];
        }
      }
      coef[0] = gg;
    }
    const double an = alpha[n], wn_ = Wm[(int64_t)n * mld + col0 + c];
    for (int q = 0; q < nk; q++) {
      const double f = 2.0 * coef[q] * dks[q];
      const double* xr = xs_tr + ((int64_t)q * npad + n) * d;
      const double* xc = sx + (q * TILE + c) * ldx;
#pragma unroll
      for (int m = 0; m < MAXD; m++)
        if (m < d) {
          const double dk = f * (xc[m] - xr[m]) * hyp.invl[q][m];
          gmu[m] = fma(an, dk, gmu[m]);
          gs[m] = fma(wn_, dk, gs[m]);
        }
    }
  }
#pragma unroll
  for (int m = 0; m < MAXD; m++)
    if (m < d) {
      sred[((rg * TILE + c) * 2 + 0) * d + m] = gmu[m];
      sred[((rg * TILE + c) * 2 + 1) * d + m] = gs[m];
    }
  __syncthreads();
  if (tid >= TILE) return;
  double g1[MAXD], g2[MAXD];
  for (int m = 0; m < d; m++) {
    g1[m] = g2[m] = 0.0;
    for (int r = 0; r < 4; r++) {
      g1[m] += sred[((r * TILE + c) * 2 + 0) * d + m];
      g2[m] += sred[((r * TILE + c) * 2 + 1) * d + m];
    }
  }
  if (nsplit > 1) {
    double* gp = gpart + ((int64_t)blockIdx.y * mld + col0 + c) * 2 * d;
    for (int m = 0; m < d; m++) {
      gp[m] = g1[m];
      gp[d + m] = g2[m];
    }
    return;
  }
  if (mg >= M) return;
  grad_epilogue(epi, d, mg, mean_add, dmean_add, g1, g2, mu_io, var_io, dmean_out, dvar_out);
}

// grid (cols / 128), one test point per thread: fixed-order sum of the row-split partials, then the epilogue.
__global__ void __launch_bounds__(128) predict_grad_finish_kernel(int d, int nsplit, const double* __restrict__ gpart,
                                                                  int mld, int64_t M, int64_t m_begin, avn_epilogue epi,
                                                                  const double* __restrict__ mean_add,
                                                                  const double* __restrict__ dmean_add,
                                                                  double* __restrict__ mu_io, double* __restrict__ var_io,
                                                                  double* __restrict__ dmean_out,
                                                                  double* __restrict__ dvar_out) {
  const int64_t col = (int64_t)blockIdx.x * 128 + threadIdx.x, mg = m_begin + col;
  if (col >= mld || mg >= M) return;
  double g1[MAXD], g2[MAXD];
  for (int m = 0; m < d; m++) g1[m] = g2[m] = 0.0;
  for (int s = 0; s < nsplit; s++) {
    const double* gp = gpart + ((int64_t)s * mld + col) * 2 * d;
    for (int m = 0; m < d; m++) {
      g1[m] += gp[m];
      g2[m] += gp[d + m];
    }
  }
  grad_epilogue(epi, d, mg, mean_add, dmean_add, g1, g2, mu_io, var_io, dmean_out, dvar_out);
}

// ------------------------------------------------------------------------------------------------
// Rank-1 extension of a factorised state by ONE training point, hyperparameters unchanged (the data
// appends of BO / inverse_opt, gpmcmc.py:881-904 / :1197-1205, between two fits):
//     L' = [[L, 0], [v^T, lam]],  v = T k,  lam^2 = (c + gv + jitter) - |v|^2
//     T' = [[T, 0], [-(T^T v)^T / lam, 1 / lam]],  alpha' = [alpha - w bn / lam ; bn / lam],  bn = (z_new - k^T alpha) / lam
// Four HBM-bound launches, T (lower triangle) read twice:
//   kvec_kernel    k = cov(X, x_new) and the block partials of k^T alpha        grid (npad / 256)
//   beta_kernel    v = T k, block partials of |v|^2                            grid (nb)      (factor.cuh)
//   alpha_kernel   w = T^T v                                                   grid (nb)
//   append_kernel  row N of T, alpha update, scaled inputs of the new row      grid (npad / 256)
// Needs N < npad (room in the padded slab).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kvec_kernel(KernDesc kd, int N, int npad, const HypS* __restrict__ hyp_g,
                                                   const double* __restrict__ xs_tr, const double* __restrict__ x2_tr,
                                                   const double* __restrict__ alpha, const double* __restrict__ xnew,
                                                   double* __restrict__ kvec, double* __restrict__ mu_part) {
  __shared__ HypS hyp;
  __shared__ double sx[MAXK][MAXD | 1];
  __shared__ double sx2[MAXK];
  __shared__ double red[32];
  const int tid = threadIdx.x, n = blockIdx.x * 256 + tid;
  for (int e = tid; e < (int)(sizeof(HypS) / sizeof(double)); e += 256)
    reinterpret_cast<double*>(&hyp)[e] = reinterpret_cast<const double*>(hyp_g)[e];
  __syncthreads();
  if (tid < kd.nkern) {
    double tmp[MAXD];
    for (int m = 0; m < kd.d; m++) {
      tmp[m] = __dmul_rn(xnew[m], hyp.invl[tid][m]);
      sx[tid][m] = tmp[m];
    }
    sx2[tid] = sumsq_numpy_order(tmp, kd.d);
  }
  __syncthreads();
  double v = 0.0;
  if (n < N)
    v = cov_fold(kd, hyp, xs_tr + (int64_t)n * kd.d, (int64_t)npad * kd.d, x2_tr + n, npad, &sx[0][0], MAXD | 1, sx2, 1);
  if (n < npad) kvec[n] = v;
  const double s = block_sum(n < N ? v * alpha[n] : 0.0, red);
  if (tid == 0) mu_part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) append_kernel(KernDesc kd, int N, int npad, const HypS* __restrict__ hyp_g,
                                                     const double* __restrict__ xnew, const double* __restrict__ znew,
                                                     const double* __restrict__ wvec, const double* __restrict__ mu_part,
                                                     const double* __restrict__ fpart, double* __restrict__ T,
                                                     double* __restrict__ alpha, double* __restrict__ xs,
                                                     double* __restrict__ x2, int32_t* __restrict__ info) {
  __shared__ HypS hyp;
  __shared__ double sx[MAXK][MAXD | 1];
  __shared__ double sx2[MAXK];
  const int tid = threadIdx.x, j = blockIdx.x * 256 + tid;
  for (int e = tid; e < (int)(sizeof(HypS) / sizeof(double)); e += 256)
    reinterpret_cast<double*>(&hyp)[e] = reinterpret_cast<const double*>(hyp_g)[e];
  __syncthreads();
  if (tid < kd.nkern) {
    double tmp[MAXD];
    for (int m = 0; m < kd.d; m++) {
      tmp[m] = __dmul_rn(xnew[m], hyp.invl[tid][m]);
      sx[tid][m] = tmp[m];
    }
    sx2[tid] = sumsq_numpy_order(tmp, kd.d);
  }
  __syncthreads();
  // diagonal entry exactly as cov_kernel builds it: full-form kernel value at r2 = 0, then + (gv + jitter)
  const double c = cov_fold(kd, hyp, &sx[0][0], MAXD | 1, sx2, 1, &sx[0][0], MAXD | 1, sx2, 1);
  double vv = 0.0, mu = 0.0;   // fixed-order sums of the block partials (every thread the same values)
  for (int k = 0; k < npad / TILE; k++) vv += fpart[2 * k];
  for (int k = 0; k < (npad + 255) / 256; k++) mu += mu_part[k];
  const double s = __dadd_rn(c, hyp.gv + kd.jitter) - vv;
  if (!(s > 0.0)) {
    if (j == 0) info[0] = N + 1;
    return;
  }
  const double lam = sqrt(s), bn = (znew[0] - mu) / lam;
  if (j < N) {
    const double w = wvec[j];
    T[(int64_t)N * npad + j] = -w / lam;
    alpha[j] -= w * bn / lam;
  } else if (j == N) {
    T[(int64_t)N * npad + N] = 1.0 / lam;
    alpha[N] = bn / lam;
    for (int k = 0; k < kd.nkern; k++) {
      for (int m = 0; m < kd.d; m++) xs[((int64_t)k * npad + N) * kd.d + m] = sx[k][m];
      x2[(int64_t)k * npad + N] = sx2[k];
    }
  }
}

}  // namespace avn

#include "kinv_fast.cuh"
#include "kinv_fold.cuh"

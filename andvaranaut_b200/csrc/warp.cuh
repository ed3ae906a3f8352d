// Composite input/output warps ("wgp") with forward-mode parameter Jacobians, evaluated per
// hyperparameter sample on the device.
//
// Semantics follow andvaranaut/transform.py: stage formulas :193-428, chain order / parameter packing /
// running zero image :431-554; the data-dependent stages (meanstd, stddev, stdshift, minshift, maxmin,
// pzero) are functions of the preceding learnable parameters exactly as in the reference's PyTensor
// mode (:448-452,:527-533 used from gpmcmc.py:224-231,:275-277), so their Jacobians carry the
// reduction terms that autodiff produces there.
#pragma once
#ifndef AVN_SKIP_DEV_HEADER
#include "avn_dev.cuh"
#endif

namespace avn {

struct WarpProgs {
  avn_warp_prog xw[MAXD];
  avn_warp_prog yw;
};

// every data-dependent / affine stage collapses to v' = a + b v with dual coefficients
struct AffineCoef {
  double a, b;
  double da[MAXWP], db[MAXWP];
};

__device__ __forceinline__ bool is_affine_family(int op) {
  return op == AVN_W_AFFINE_CONST || op == AVN_W_AFFINE || (op >= AVN_W_STDSHIFT && op <= AVN_W_PZERO);
}

// Non-affine stages.  p: stage parameters (learnable values or frozen constants); pidx: index of the
// first learnable parameter in the dual vector or -1 when frozen.  Updates v, dv[0..np), and adds
// log(d con / d v) and its parameter derivatives into lder / dlder when track != 0.
__device__ __forceinline__ void apply_nonaffine(int op, int pidx, const double* __restrict__ p, int np, double& v,
                                                double* __restrict__ dv, int track, double& lder,
                                                double* __restrict__ dlder) {
  switch (op) {
    case AVN_W_LOG: {
      double iv = 1.0 / v;
      if (track) {
        lder += log(iv);
        for (int q = 0; q < np; q++) dlder[q] -= dv[q] * iv;
      }
      for (int q = 0; q < np; q++) dv[q] *= iv;
      v = log(v);
    } break;
    case AVN_W_ARCSINH: {
      const double a = p[0], b = p[1], c = p[2], d = p[3];
      double t = (v - c) / d;
      double s = asinh(t);
      double rt = rsqrt(1.0 + t * t);  // ds/dt
      double q2 = d * d + (v - c) * (v - c);
      if (track) {
        lder += log(b / sqrt(q2));
        for (int q = 0; q < np; q++) {
          double dvc = dv[q] - ((pidx >= 0 && q == pidx + 2) ? 1.0 : 0.0);
          double dd = (pidx >= 0 && q == pidx + 3) ? 1.0 : 0.0;
          dlder[q] += ((pidx >= 0 && q == pidx + 1) ? 1.0 / b : 0.0) - (d * dd + (v - c) * dvc) / q2;
        }
      }
      for (int q = 0; q < np; q++) {
        double dvc = dv[q] - ((pidx >= 0 && q == pidx + 2) ? 1.0 : 0.0);
        double dd = (pidx >= 0 && q == pidx + 3) ? 1.0 : 0.0;
        double dt = (dvc - t * dd) / d;
        double r = b * rt * dt;
        if (pidx >= 0 && q == pidx) r += 1.0;
        if (pidx >= 0 && q == pidx + 1) r += s;
        dv[q] = r;
      }
      v = a + b * s;
    } break;
    case AVN_W_BOXCOX:
    case AVN_W_BOXCOX_CONST: {
      const double lam = p[0];
      const double qq = lam + 1.0;
      double av = fabs(v);
      double sg = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
      double pw = pow(av, qq);
      double lav = (av > 0.0) ? log(av) : 0.0;
      double vn = (sg * pw - 1.0) / qq;
      double dvv = pow(av, lam);  // d con / d v
      const bool learn = (op == AVN_W_BOXCOX) && pidx >= 0;
      if (track) {
        lder += log(dvv);
        for (int q = 0; q < np; q++) dlder[q] += ((learn && q == pidx) ? lav : 0.0) + lam * dv[q] / v;
      }
      for (int q = 0; q < np; q++) {
        double r = dvv * dv[q];
        if (learn && q == pidx) r += sg * pw * lav / qq - vn / qq;
        dv[q] = r;
      }
      v = vn;
    } break;
    case AVN_W_SINHARCSINH:
    case AVN_W_SAL: {
      const double a = p[0], b = p[1];
      const bool is_sal = (op == AVN_W_SAL);
      const double c = is_sal ? p[2] : 0.0, d = is_sal ? p[3] : 1.0;
      double s = asinh(v);
      double u = b * s - a;
      double sh = sinh(u), ch = cosh(u);
      double i1 = 1.0 / (1.0 + v * v);
      double rt = sqrt(i1);
      double th = sh / ch;
      if (track) lder += log(b * d * ch * rt);
      for (int q = 0; q < np; q++) {
        double du = b * dv[q] * rt;
        if (pidx >= 0 && q == pidx + 1) du += s;
        if (pidx >= 0 && q == pidx) du -= 1.0;
        if (track) {
          double r = th * du - v * dv[q] * i1;
          if (pidx >= 0 && q == pidx + 1) r += 1.0 / b;
          if (is_sal && pidx >= 0 && q == pidx + 3) r += 1.0 / d;
          dlder[q] += r;
        }
        double r = d * ch * du;
        if (is_sal && pidx >= 0 && q == pidx + 2) r += 1.0;
        if (is_sal && pidx >= 0 && q == pidx + 3) r += sh;
        dv[q] = r;
      }
      v = is_sal ? (c + d * sh) : sh;
    } break;
    case AVN_W_KUMARASWAMY: {
      const double a = p[0], b = p[1];
      double xa = pow(v, a);
      double w = 1.0 - xa;
      double wb = pow(w, b);
      double lv = (v > 0.0) ? log(v) : 0.0;
      double lw = (w > 0.0) ? log(w) : 0.0;
      double xam1 = pow(v, a - 1.0);
      double wbm1 = pow(w, b - 1.0);
      if (track) lder += log(a * b * xam1 * wbm1);
      for (int q = 0; q < np; q++) {
        double dxa = a * xam1 * dv[q];
        if (pidx >= 0 && q == pidx) dxa += xa * lv;
        double dw = -dxa;
        double dwb = b * wbm1 * dw;
        if (pidx >= 0 && q == pidx + 1) dwb += wb * lw;
        if (track) {
          double r = (a - 1.0) * dv[q] / v + (b - 1.0) * dw / w;
          if (pidx >= 0 && q == pidx) r += 1.0 / a + lv;
          if (pidx >= 0 && q == pidx + 1) r += 1.0 / b + lw;
          dlder[q] += r;
        }
        dv[q] = -dwb;
      }
      v = 1.0 - wb;
    } break;
    default:
      break;
  }
}

__device__ __forceinline__ void apply_affine(const AffineCoef& c, int np, double& v, double* __restrict__ dv,
                                             int track, double& lder, double* __restrict__ dlder) {
  for (int q = 0; q < np; q++) dv[q] = c.da[q] + c.db[q] * v + c.b * dv[q];
  v = __dadd_rn(c.a, __dmul_rn(c.b, v));
  if (track) {
    lder += log(c.b);
    for (int q = 0; q < np; q++) dlder[q] += c.db[q] / c.b;
  }
}

// Frozen forward / inverse of a single stage (all coefficients constant): used on test points.
__device__ __forceinline__ double stage_rev_const(const avn_warp_stage& st, double z) {
  const double* p = st.c;
  switch (st.op) {
    case AVN_W_AFFINE_CONST:
    case AVN_W_AFFINE:
      return (z - p[0]) / p[1];
    case AVN_W_LOG:
      return exp(z);
    case AVN_W_ARCSINH:
      return p[2] + p[3] * sinh((z - p[0]) / p[1]);
    case AVN_W_BOXCOX:
    case AVN_W_BOXCOX_CONST: {
      double q = p[0] + 1.0;
      double t = z * q + 1.0;
      double sg = (t > 0.0) ? 1.0 : ((t < 0.0) ? -1.0 : 0.0);
      return sg * pow(fabs(t), 1.0 / q);
    }
    case AVN_W_SINHARCSINH:
      return sinh((asinh(z) + p[0]) / p[1]);
    case AVN_W_SAL:
      return sinh((asinh((z - p[2]) / p[3]) + p[0]) / p[1]);
    case AVN_W_KUMARASWAMY:
      return pow(1.0 - pow(1.0 - z, 1.0 / p[1]), 1.0 / p[0]);
    default:
      return z;
  }
}

__device__ __forceinline__ double prog_rev_const(const avn_warp_prog& pr, double z) {
  for (int s = pr.nstages - 1; s >= 0; s--) z = stage_rev_const(pr.st[s], z);
  return z;
}

// Frozen inverse of a single stage together with its derivative d rev / d z (chain rule of the BO refine graph,
// gpmcmc.py:781-783: revmc inside a differentiated PyTensor expression).
__device__ __forceinline__ double stage_rev_const_d(const avn_warp_stage& st, double z, double& dz) {
  const double* p = st.c;
  switch (st.op) {
    case AVN_W_AFFINE_CONST:
    case AVN_W_AFFINE:
      dz = 1.0 / p[1];
      return (z - p[0]) / p[1];
    case AVN_W_LOG: {
      const double e = exp(z);
      dz = e;
      return e;
    }
    case AVN_W_ARCSINH: {
      const double u = (z - p[0]) / p[1];
      dz = p[3] * cosh(u) / p[1];
      return p[2] + p[3] * sinh(u);
    }
    case AVN_W_BOXCOX:
    case AVN_W_BOXCOX_CONST: {
      const double q = p[0] + 1.0;
      const double t = z * q + 1.0;
      const double sg = (t > 0.0) ? 1.0 : ((t < 0.0) ? -1.0 : 0.0);
      dz = pow(fabs(t), 1.0 / q - 1.0);
      return sg * pow(fabs(t), 1.0 / q);
    }
    case AVN_W_SINHARCSINH: {
      const double u = (asinh(z) + p[0]) / p[1];
      dz = cosh(u) / (p[1] * sqrt(1.0 + z * z));
      return sinh(u);
    }
    case AVN_W_SAL: {
      const double w = (z - p[2]) / p[3];
      const double u = (asinh(w) + p[0]) / p[1];
      dz = cosh(u) / (p[1] * p[3] * sqrt(1.0 + w * w));
      return sinh(u);
    }
    case AVN_W_KUMARASWAMY: {
      const double s = pow(1.0 - z, 1.0 / p[1]);
      dz = (1.0 / p[0]) * pow(1.0 - s, 1.0 / p[0] - 1.0) * (1.0 / p[1]) * pow(1.0 - z, 1.0 / p[1] - 1.0);
      return pow(1.0 - s, 1.0 / p[0]);
    }
    default:
      dz = 1.0;
      return z;
  }
}

__device__ __forceinline__ double prog_rev_const_d(const avn_warp_prog& pr, double z, double& dz) {
  double d = 1.0;
  for (int s = pr.nstages - 1; s >= 0; s--) {
    double ds;
    z = stage_rev_const_d(pr.st[s], z, ds);
    d *= ds;
  }
  dz = d;
  return z;
}

// Run one composite warp over a strided column of N values, in place.
//   val[n*vstride]                 running value (in: raw data, out: converted)
//   dual[n*dstride + q], q < np    d value / d param_q (zeroed here)
//   pvals                          the warp's learnable parameters (np of them)
// When track != 0 the per-thread sums of log(d con/d y) and their parameter derivatives are
// accumulated into lsum / dlsum (to be block-reduced by the caller).
// NOTE: val / dual / sh carry data between threads across __syncthreads(); they must NOT be
// __restrict__ (noalias lets the compiler hoist their loads above the barrier).
// sh: shared scratch, >= 128 + sizeof(AffineCoef)/8 + 2*(MAXWP+1) doubles (192 is enough); up to 32 warps per block.
__device__ void run_warp_column(const avn_warp_prog& pr, const double* __restrict__ pvals, int N,
                                double* val, int64_t vstride, double* dual,
                                int64_t dstride, int track, double& lsum, double* __restrict__ dlsum,
                                double* sh) {
  const int np = pr.nparams;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* red = sh;                                   // 32 doubles: block_sum scratch
  double* stat = sh + 32;                             // 96 doubles: per-warp min / max and their indices
  AffineCoef* coef = reinterpret_cast<AffineCoef*>(sh + 128);
  double* zero = sh + 128 + (sizeof(AffineCoef) + 7) / 8;  // running image of 0: value + np duals
  for (int n = tid; n < N; n += nt)
    for (int q = 0; q < np; q++) dual[n * dstride + q] = 0.0;
  if (tid == 0) {
    zero[0] = 0.0;
    for (int q = 0; q < np; q++) zero[1 + q] = 0.0;
  }
  __syncthreads();
  for (int s = 0; s < pr.nstages; s++) {
    const avn_warp_stage st = pr.st[s];
    const int op = st.op;
    double p[4];
    for (int i = 0; i < 4; i++) p[i] = (st.pidx >= 0 && st.pidx + i < np) ? pvals[st.pidx + i] : st.c[i];
    if (is_affine_family(op)) {
      // ---- statistics of the running data (with duals) ----
      double mean = 0, sd = 1, dmean[MAXWP], dsd[MAXWP];
      double vmin = 0, vmax = 0, dmin[MAXWP], dmax[MAXWP];
      for (int q = 0; q < MAXWP; q++) dmean[q] = dsd[q] = dmin[q] = dmax[q] = 0.0;
      const bool need_std = (op == AVN_W_STDSHIFT || op == AVN_W_MEANSTD || op == AVN_W_STDDEV || op == AVN_W_PZERO);
      const bool need_mm = (op == AVN_W_MINSHIFT || op == AVN_W_MAXMIN);
      if (need_std) {
        double sv = 0;
        for (int n = tid; n < N; n += nt) sv += val[n * vstride];
        mean = block_sum(sv, red) / N;
        for (int q = 0; q < np; q++) {
          double sq = 0;
          for (int n = tid; n < N; n += nt) sq += dual[n * dstride + q];
          dmean[q] = block_sum(sq, red) / N;
        }
        double s2 = 0;
        for (int n = tid; n < N; n += nt) {
          double c = val[n * vstride] - mean;
          s2 += c * c;
        }
        sd = sqrt(block_sum(s2, red) / N);
        for (int q = 0; q < np; q++) {
          double sq = 0;
          for (int n = tid; n < N; n += nt) sq += (val[n * vstride] - mean) * (dual[n * dstride + q] - dmean[q]);
          dsd[q] = block_sum(sq, red) / (N * sd);
        }
      }
      if (need_mm) {
        // argmin / argmax with first-index tie break (np.argmin semantics)
        double lo = INFINITY, hi = -INFINITY;
        int ilo = 0x7fffffff, ihi = 0x7fffffff;
        for (int n = tid; n < N; n += nt) {
          double x = val[n * vstride];
          if (x < lo) { lo = x; ilo = n; }
          if (x > hi) { hi = x; ihi = n; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          double lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
          int il2 = __shfl_xor_sync(0xffffffffu, ilo, o), ih2 = __shfl_xor_sync(0xffffffffu, ihi, o);
          if (lo2 < lo || (lo2 == lo && il2 < ilo)) { lo = lo2; ilo = il2; }
          if (hi2 > hi || (hi2 == hi && ih2 < ihi)) { hi = hi2; ihi = ih2; }
        }
        double* slo = stat;
        double* shi = stat + 32;
        int* silo = reinterpret_cast<int*>(stat + 64);
        int* sihi = silo + 32;
        __syncthreads();
        if ((tid & 31) == 0) {
          slo[tid >> 5] = lo; shi[tid >> 5] = hi; silo[tid >> 5] = ilo; sihi[tid >> 5] = ihi;
        }
        __syncthreads();
        lo = slo[0]; hi = shi[0]; ilo = silo[0]; ihi = sihi[0];
        for (int w = 1; w < (nt >> 5); w++) {
          if (slo[w] < lo || (slo[w] == lo && silo[w] < ilo)) { lo = slo[w]; ilo = silo[w]; }
          if (shi[w] > hi || (shi[w] == hi && sihi[w] < ihi)) { hi = shi[w]; ihi = sihi[w]; }
        }
        vmin = lo; vmax = hi;
        for (int q = 0; q < np; q++) {
          dmin[q] = dual[ilo * dstride + q];
          dmax[q] = dual[ihi * dstride + q];
        }
        __syncthreads();
      }
      if (tid == 0) {
        AffineCoef c;
        for (int q = 0; q < MAXWP; q++) c.da[q] = c.db[q] = 0.0;
        switch (op) {
          case AVN_W_AFFINE_CONST:
            c.a = st.c[0]; c.b = st.c[1];
            break;
          case AVN_W_AFFINE:
            c.a = p[0]; c.b = p[1];
            if (st.pidx >= 0) { c.da[st.pidx] = 1.0; c.db[st.pidx + 1] = 1.0; }
            break;
          case AVN_W_STDSHIFT:
            c.a = p[0]; c.b = 1.0 / sd;
            if (st.pidx >= 0) c.da[st.pidx] = 1.0;
            for (int q = 0; q < np; q++) c.db[q] = -dsd[q] / (sd * sd);
            break;
          case AVN_W_MEANSTD:
            c.a = -mean / sd; c.b = 1.0 / sd;
            for (int q = 0; q < np; q++) {
              c.da[q] = -dmean[q] / sd + mean * dsd[q] / (sd * sd);
              c.db[q] = -dsd[q] / (sd * sd);
            }
            break;
          case AVN_W_MINSHIFT:
            c.a = -vmin * 1000.0; c.b = 1.0;
            for (int q = 0; q < np; q++) c.da[q] = -1000.0 * dmin[q];
            break;
          case AVN_W_STDDEV:
            c.a = 0.0; c.b = 1.0 / sd;
            for (int q = 0; q < np; q++) c.db[q] = -dsd[q] / (sd * sd);
            break;
          case AVN_W_MAXMIN: {
            const double safety = 0.01;
            double xm = (vmax - vmin) / (1.0 - 2.0 * safety);
            c.a = -vmin / xm + safety; c.b = 1.0 / xm;
            for (int q = 0; q < np; q++) {
              double dxm = (dmax[q] - dmin[q]) / (1.0 - 2.0 * safety);
              c.da[q] = -dmin[q] / xm + vmin * dxm / (xm * xm);
              c.db[q] = -dxm / (xm * xm);
            }
          } break;
          default: {  // AVN_W_PZERO
            c.a = -zero[0] / sd; c.b = 1.0 / sd;
            for (int q = 0; q < np; q++) {
              c.da[q] = -zero[1 + q] / sd + zero[0] * dsd[q] / (sd * sd);
              c.db[q] = -dsd[q] / (sd * sd);
            }
          } break;
        }
        *coef = c;
      }
      __syncthreads();
      const AffineCoef c = *coef;
      for (int n = tid; n < N; n += nt) {
        double v = val[n * vstride];
        double dv[MAXWP];
        for (int q = 0; q < np; q++) dv[q] = dual[n * dstride + q];
        apply_affine(c, np, v, dv, track, lsum, dlsum);
        val[n * vstride] = v;
        for (int q = 0; q < np; q++) dual[n * dstride + q] = dv[q];
      }
      if (tid == 0) {
        double dummy = 0, ddummy[MAXWP];
        apply_affine(c, np, zero[0], zero + 1, 0, dummy, ddummy);
      }
    } else {
      for (int n = tid; n < N; n += nt) {
        double v = val[n * vstride];
        double dv[MAXWP];
        for (int q = 0; q < np; q++) dv[q] = dual[n * dstride + q];
        apply_nonaffine(op, st.pidx, p, np, v, dv, track, lsum, dlsum);
        val[n * vstride] = v;
        for (int q = 0; q < np; q++) dual[n * dstride + q] = dv[q];
      }
      if (tid == 0) {
        double dummy = 0, ddummy[MAXWP];
        apply_nonaffine(op, st.pidx, p, np, zero[0], zero + 1, 0, dummy, ddummy);
      }
    }
    __syncthreads();
  }
}

}  // namespace avn

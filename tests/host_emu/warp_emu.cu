// Host emulation of csrc/warp.cuh with a one-thread "block" (tid 0 of 1): checks the per-stage
// warp math and Jacobians on the CPU build box (no GPU).  Built and driven by tests/test_warp_emu.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define AVN_HOST_EMU 1
struct Dim3 { int x; };
static Dim3 threadIdx = {0}, blockDim = {1};
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __syncthreads() ((void)0)
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }  // single lane: only used for idempotent min/max merges
static inline int __shfl_xor_sync(unsigned, int v, int) { return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::fma;
#include "../../include/avn_gp.h"
namespace avn {
constexpr int MAXD = AVN_MAX_D, MAXK = AVN_MAX_KERN, MAXWP = AVN_MAX_WPARAMS;
static inline double block_sum(double v, double*) { return v; }
}
#define AVN_SKIP_DEV_HEADER 1
#include "../../andvaranaut_b200/csrc/warp.cuh"

extern "C" int warp_emu_run(const avn_warp_prog* pr, const double* pvals, int N, double* val, double* dual,
                            int track, double* lsum, double* dlsum) {
  static double sh[256];
  double ls = 0, dls[AVN_MAX_WPARAMS];
  for (int q = 0; q < AVN_MAX_WPARAMS; q++) dls[q] = 0;
  avn::run_warp_column(*pr, pvals, N, val, 1, dual, AVN_MAX_WPARAMS, track, ls, dls, sh);
  *lsum = ls;
  for (int q = 0; q < AVN_MAX_WPARAMS; q++) dlsum[q] = dls[q];
  return 0;
}
extern "C" double warp_emu_rev(const avn_warp_prog* pr, double z) { return avn::prog_rev_const(*pr, z); }

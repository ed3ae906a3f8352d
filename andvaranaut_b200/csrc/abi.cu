// C ABI of libavn_gp.so: argument checking, workspace layout and the launch sequences.
// See include/avn_gp.h for the contract and the reference call sites each entry point replaces.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>

#include "kernels.cuh"

using namespace avn;

static thread_local std::string g_err;
static int fail(const char* msg) {
  g_err = msg;
  return -1;
}
static int fail_cuda(const char* where, cudaError_t e) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return -2;
}

struct avn_gp {
  avn_model_desc desc;
  KernDesc kd;
  WarpProgs progs;
  const double* X = nullptr;
  const double* y = nullptr;
  int64_t N = 0;
  int64_t launches = 0;
  bool has_xwarp = false;
  bool profiling = false;
  int device = -1;                    // CUDA device the handle is bound to (current device at avn_gp_create, or at the
                                      // first call that enqueues work when no device was visible at create time)
  bool dev_ready = false;             // shared-memory opt-ins / occupancy of this handle's kernels done on `device`
  int sm_count = 0;                   // SMs of `device`
  int fac_resident = 0;               // CTAs of the persistent factor kernel that are resident at once on `device`
  int fac_resident_fused = 0;         // the same with the fused-panel shared-memory footprint (three tiles)
  unsigned max_spins = 1u << 26;      // bound of the factor kernel's flag waits, in polls (avn_gp_set_debug)
  int fault = 0;                      // fault injection for tests (avn_gp_set_debug)
  int max_groups = 1;                 // independent sample groups on internal streams (avn_gp_set_streams)
  cudaStream_t gstream[8] = {};
  cudaEvent_t gev[9] = {};            // [0]: fork point on the caller's stream, [1+g]: join of group g
  cudaEvent_t ev[2 * AVN_PH_COUNT] = {};
  bool ev_used[AVN_PH_COUNT] = {};
  double acc_ms[AVN_PH_COUNT] = {};
  // few samples: the output-warp column of warp_kernel runs on its own stream beside the input columns, scale and cov
  cudaStream_t ystream = nullptr;
  cudaEvent_t yev_fork = nullptr, yev_join = nullptr;
  // avn_gp_loglik_grad_host: the whole host-to-host evaluation captured once as a CUDA graph and replayed
  cudaStream_t hstream = nullptr;     // library-owned: stream capture is not allowed on the legacy default stream
  cudaEvent_t hev = nullptr;          // orders the replay behind the caller's stream
  cudaGraphExec_t hexec = nullptr;
  int64_t hlaunches = 0;              // kernel launches inside the captured graph
  struct HostKey {
    const void* theta_host; const void* out_host; const void* staging; const void* ws; const void* X; const void* y;
    int64_t B, N; int want_grad; unsigned max_spins; int fault;
    bool operator==(const HostKey& o) const {
      return theta_host == o.theta_host && out_host == o.out_host && staging == o.staging && ws == o.ws && X == o.X &&
             y == o.y && B == o.B && N == o.N && want_grad == o.want_grad && max_spins == o.max_spins && fault == o.fault;
    }
  } hkey = {};
};

static void host_graph_drop(avn_gp* gp) {
  if (gp->hexec) cudaGraphExecDestroy(gp->hexec);
  gp->hexec = nullptr;
}

// RAII phase marker: records start/stop events when profiling is on
struct Phase {
  avn_gp* gp;
  int id;
  cudaStream_t st;
  Phase(avn_gp* g, int i, cudaStream_t s) : gp(g), id(i), st(s) {
    if (gp->profiling) cudaEventRecord(gp->ev[2 * id], st);
  }
  ~Phase() {
    if (gp->profiling) {
      cudaEventRecord(gp->ev[2 * id + 1], st);
      gp->ev_used[id] = true;
    }
  }
};
// Every entry point that enqueues work runs with the handle's device current and restores the caller's on exit, so a
// handle may be used from any host thread (new threads start on device 0) and several handles on several devices may
// share one process.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DevGuard(avn_gp* gp) {
    err = cudaGetDevice(&prev);
    if (err != cudaSuccess) return;
    if (gp->device < 0) gp->device = prev;
    if (gp->device != prev) {
      err = cudaSetDevice(gp->device);
      switched = err == cudaSuccess;
    }
  }
  ~DevGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
static int ensure_ready(avn_gp* gp);
#define ENTER_DEVICE(gp)                                              \
  DevGuard dev_guard__(gp);                                           \
  if (dev_guard__.err != cudaSuccess) return fail_cuda("device", dev_guard__.err); \
  if (!(gp)->dev_ready) {                                             \
    int rc__ = ensure_ready(gp);                                      \
    if (rc__) return rc__;                                            \
  }

static void phases_reset(avn_gp* gp) {
  for (int i = 0; i < AVN_PH_COUNT; i++) gp->ev_used[i] = false;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int64_t npad_of(int64_t N) { return align_up(N < 1 ? 1 : N, TILE); }

extern "C" const char* avn_last_error(void) { return g_err.c_str(); }
extern "C" int avn_version(void) { return 110; }   // 110: avn_gp_loglik_grad_host / avn_gp_host_wait / avn_gp_host_staging_bytes

extern "C" int avn_gp_create(const avn_model_desc* desc, avn_gp** out) {
  if (!desc || !out) return fail("avn_gp_create: null argument");
  if (desc->d < 1 || desc->d > AVN_MAX_D) return fail("avn_gp_create: d out of range [1,16]");
  if (desc->nkern < 1 || desc->nkern > AVN_MAX_KERN) return fail("avn_gp_create: nkern out of range [1,4]");
  int nrq = 0;
  for (int k = 0; k < desc->nkern; k++) {
    if (desc->kern[k] < AVN_RBF || desc->kern[k] > AVN_RATQUAD) return fail("avn_gp_create: unknown kernel id");
    if (desc->kern[k] == AVN_RATQUAD) nrq++;
    if (k > 0 && desc->op[k - 1] != AVN_ADD && desc->op[k - 1] != AVN_MUL) return fail("avn_gp_create: bad kernel op");
  }
  if (nrq > 1) return fail("avn_gp_create: at most one RatQuad kernel (gpmcmc.py:287)");
  avn_gp* gp = new avn_gp();
  gp->desc = *desc;
  KernDesc& kd = gp->kd;
  memset(&kd, 0, sizeof(kd));
  kd.d = desc->d;
  kd.nkern = desc->nkern;
  kd.noise = desc->noise ? 1 : 0;
  kd.has_alpha = nrq;
  kd.jitter = desc->jitter;
  for (int k = 0; k < desc->nkern; k++) {
    kd.kern[k] = desc->kern[k];
    kd.op[k] = desc->op[k];
  }
  int n_iw = 0;
  for (int m = 0; m < desc->d; m++) {
    const avn_warp_prog& p = desc->xwarp[m];
    if (p.nstages < 0 || p.nstages > AVN_MAX_STAGES || p.nparams < 0 || p.nparams > AVN_MAX_WPARAMS) {
      delete gp;
      return fail("avn_gp_create: x warp program out of range");
    }
    gp->progs.xw[m] = p;
    if (p.nstages > 0) {
      gp->has_xwarp = true;
      n_iw += p.nparams;
    }
  }
  for (int m = desc->d; m < AVN_MAX_D; m++) memset(&gp->progs.xw[m], 0, sizeof(avn_warp_prog));
  const avn_warp_prog& yp = desc->ywarp;
  if (yp.nstages < 0 || yp.nstages > AVN_MAX_STAGES || yp.nparams < 0 || yp.nparams > AVN_MAX_WPARAMS) {
    delete gp;
    return fail("avn_gp_create: y warp program out of range");
  }
  gp->progs.yw = yp;
  int p = 0;
  kd.off_gv = p;
  if (kd.noise) p += 1;
  kd.off_l = p;
  p += kd.d * kd.nkern;
  kd.off_kv = p;
  p += kd.nkern;
  kd.off_iw = p;
  kd.n_iw = n_iw;
  p += n_iw;
  kd.off_cw = p;
  kd.n_cw = yp.nstages > 0 ? yp.nparams : 0;
  p += kd.n_cw;
  kd.off_alpha = p;
  if (kd.has_alpha) p += 1;
  kd.P = p;
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess) gp->device = dev;   // no device visible (CPU-only symbol checks): bound later
  else (void)cudaGetLastError();
  *out = gp;
  return 0;
}

extern "C" void avn_gp_destroy(avn_gp* gp) {
  if (!gp) return;
  DevGuard guard(gp);
  for (auto& e : gp->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : gp->gev)
    if (e) cudaEventDestroy(e);
  for (auto& s : gp->gstream)
    if (s) cudaStreamDestroy(s);
  host_graph_drop(gp);
  if (gp->yev_fork) cudaEventDestroy(gp->yev_fork);
  if (gp->yev_join) cudaEventDestroy(gp->yev_join);
  if (gp->ystream) cudaStreamDestroy(gp->ystream);
  if (gp->hev) cudaEventDestroy(gp->hev);
  if (gp->hstream) cudaStreamDestroy(gp->hstream);
  delete gp;
}

extern "C" int avn_gp_set_debug(avn_gp* gp, int wait_bound_log2, int fault) {
  if (!gp) return fail("avn_gp_set_debug: null handle");
  if (wait_bound_log2 < 10 || wait_bound_log2 > 31) return fail("avn_gp_set_debug: wait_bound_log2 out of range [10,31]");
  if (fault < 0 || fault > 1) return fail("avn_gp_set_debug: unknown fault");
  gp->max_spins = wait_bound_log2 >= 31 ? 0x7fffffffu : (1u << wait_bound_log2);
  gp->fault = fault;
  return 0;
}

extern "C" int avn_gp_set_profiling(avn_gp* gp, int enable) {
  if (!gp) return fail("avn_gp_set_profiling: null handle");
  DevGuard guard(gp);
  if (enable && !gp->ev[0]) {
    for (auto& e : gp->ev) {
      cudaError_t err = cudaEventCreate(&e);
      if (err != cudaSuccess) return fail_cuda("cudaEventCreate", err);
    }
  }
  gp->profiling = enable != 0;
  return 0;
}

extern "C" int avn_gp_phase_ms(avn_gp* gp, double* out_ms) {
  if (!gp || !out_ms) return fail("avn_gp_phase_ms: null argument");
  DevGuard guard(gp);
  for (int i = 0; i < AVN_PH_COUNT; i++) {
    out_ms[i] = 0.0;
    if (gp->profiling && gp->ev_used[i]) {
      cudaError_t err = cudaEventSynchronize(gp->ev[2 * i + 1]);
      if (err != cudaSuccess) return fail_cuda("cudaEventSynchronize", err);
      float ms = 0.f;
      err = cudaEventElapsedTime(&ms, gp->ev[2 * i], gp->ev[2 * i + 1]);
      if (err != cudaSuccess) return fail_cuda("cudaEventElapsedTime", err);
      out_ms[i] = ms;
    }
  }
  return 0;
}
extern "C" int avn_gp_num_params(const avn_gp* gp) { return gp ? gp->kd.P : -1; }
extern "C" int64_t avn_gp_last_launch_count(const avn_gp* gp) { return gp ? gp->launches : -1; }

extern "C" int avn_gp_set_data(avn_gp* gp, const double* X_dev, const double* y_dev, int64_t N) {
  if (!gp || !X_dev || !y_dev) return fail("avn_gp_set_data: null argument");
  if (N < 1 || N > 65536) return fail("avn_gp_set_data: N out of range [1,65536]");
  gp->X = X_dev;
  gp->y = y_dev;
  gp->N = N;
  return 0;
}

static const int64_t kKinvSchedInts = 258 + 3 * 256;      // scheduler words of kinv_grad_fast_single_kernel (kinv_fast.cuh)

static void layout(const avn_gp* gp, int64_t B, avn_ws_layout* L) {
  const KernDesc& kd = gp->kd;
  const int64_t npad = npad_of(gp->N), nb = npad / TILE, ntiles = nb * (nb + 1) / 2;
  int64_t off = 0;
  auto take = [&](int64_t doubles) {
    int64_t o = off;
    off += align_up(doubles * 8, 256);
    return o;
  };
  L->npad = npad;
  L->nb = nb;
  L->xw = take(B * npad * kd.d);
  L->dxw = take(gp->has_xwarp ? B * npad * kd.d * MAXWP : 0);
  L->xs = take(B * kd.nkern * npad * kd.d);
  L->x2 = take(B * kd.nkern * npad);
  L->z = take(B * npad);
  L->dz = take(kd.n_cw > 0 ? B * npad * MAXWP : 0);
  L->wstat = take(B * WSTAT);
  L->kl = take(B * npad * npad);
  L->t = take(B * npad * npad);
  L->beta = take(B * npad);
  L->alpha = take(B * npad);
  L->gpart = take(B * ntiles * MAXACC);
  L->gxpart = take(gp->has_xwarp ? B * nb * npad * kd.d : 0);
  L->fpart = take(B * nb * 2);
  // int32 progress flags of the factor kernel: lflag [B][nb], tflag [B][nb], 8 control words per stream group, sflag [B][nb], dflag [B][nb],
  // and the scheduler words of the single-sample gradient kernel (kinv_grad_fast_single_kernel)
  L->fflags = take((4 * B * nb + 8 * 8 + kKinvSchedInts + 1) / 2);
  L->total = off;
}

static WsPtrs ws_ptrs(const avn_ws_layout& L, void* ws, int64_t B) {
  char* base = static_cast<char*>(ws);
  auto at = [&](int64_t o) { return reinterpret_cast<double*>(base + o); };
  WsPtrs p;
  p.xw = at(L.xw); p.dxw = at(L.dxw); p.xs = at(L.xs); p.x2 = at(L.x2); p.z = at(L.z); p.dz = at(L.dz);
  p.wstat = at(L.wstat); p.kl = at(L.kl); p.t = at(L.t); p.beta = at(L.beta); p.alpha = at(L.alpha);
  p.gpart = at(L.gpart); p.gxpart = at(L.gxpart); p.fpart = at(L.fpart);
  p.lflag = reinterpret_cast<int32_t*>(base + L.fflags);
  p.tflag = p.lflag + B * L.nb;
  p.ctl = p.tflag + B * L.nb;
  p.sflag = p.ctl + 64;
  p.dflag = p.sflag + B * L.nb;
  p.ksched = p.dflag + B * L.nb;
  return p;
}

extern "C" size_t avn_gp_workspace_bytes(const avn_gp* gp, int64_t B) {
  if (!gp || gp->N < 1 || B < 1) return 0;
  avn_ws_layout L;
  layout(gp, B, &L);
  return (size_t)L.total;
}

extern "C" int avn_gp_workspace_layout(const avn_gp* gp, int64_t B, avn_ws_layout* out) {
  if (!gp || !out || gp->N < 1 || B < 1) return fail("avn_gp_workspace_layout: bad argument");
  layout(gp, B, out);
  return 0;
}

static const int64_t kWarpStageMaxBytes = 160 * 1024;   // shared-memory staging of a warped column (warp_kernel)

template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  // ask for the full shared-memory carveout so that several CTAs fit on one SM
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

#define LAUNCH_CHECK(name)                                  \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return fail_cuda(name, e__);    \
    gp->launches++;                                         \
  } while (0)

// dynamic shared memory of the kernels whose request depends on the model (d, nkern)
static size_t cov_smem_bytes(const KernDesc& kd) { return (size_t)(2 * kd.nkern * TILE * kd.d + 2 * kd.nkern * TILE) * 8; }
static size_t kinv_fast_smem_bytes(const KernDesc& kd) {
  const KinvFastLayout lay(kd.d);
  const size_t smem = (size_t)lay.total * 8;
  return smem < KinvG2::SMEM_BYTES ? KinvG2::SMEM_BYTES : smem;
}
static size_t kinv_fold_smem_bytes(const KernDesc& kd) {
  const KinvFoldLayout lay(kd.d);
  const size_t smem = (size_t)lay.total * 8;
  return smem < KinvG2::SMEM_BYTES ? KinvG2::SMEM_BYTES : smem;
}
static size_t kxs_smem_bytes(const KernDesc& kd) { return (size_t)(kd.nkern * TILE * (kd.d | 1) + kd.nkern * TILE) * 8; }
static size_t predict_grad_smem_bytes(const KernDesc& kd) { return kxs_smem_bytes(kd) + (size_t)(4 * TILE * 2 * kd.d) * 8; }

static int launch_kinv_fast(bool optin_only, int kind, bool gx, dim3 grid, size_t smem, cudaStream_t st, const KernDesc& kd,
                            int N, int npad, const double* theta, const WsPtrs& W, int single = 0, int nsm = 0);

// Once per handle, on the handle's device (the caller holds a DevGuard): every shared-memory opt-in this model's
// kernels need and the resident CTA count of the persistent factor kernel.  cudaFuncSetAttribute is per device, so the
// result lives in the handle -- no process-wide caches -- and nothing of this is left on the per-call path.
static int ensure_ready(avn_gp* gp) {
  const KernDesc& kd = gp->kd;
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  if (cov_smem_bytes(kd) > 48 * 1024) chk(opt_in_smem(cov_kernel, cov_smem_bytes(kd)));
  chk(opt_in_smem(factor_kernel<false>, FAC_SMEM_BYTES));
  chk(opt_in_smem(factor_kernel<true>, FAC_SMEM_BYTES_FUSED));
  chk(cudaFuncSetAttribute(warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpStageMaxBytes));
  if (kd.nkern == 1) {
    WsPtrs none{};
    int rc = launch_kinv_fast(true, kd.kern[0], gp->has_xwarp, dim3(1), kinv_fast_smem_bytes(kd), nullptr, kd, 0, 0, nullptr, none);
    if (rc) return rc;
  } else if (kd.nkern == 2 && !(kd.kern[0] == AVN_RATQUAD && kd.kern[1] == AVN_RATQUAD)) {
    const size_t smem = kinv_fold_smem_bytes(kd);
    chk(opt_in_smem(kinv_grad_fold2_kernel<false, false>, smem));
    chk(opt_in_smem(kinv_grad_fold2_kernel<true, false>, smem));
    chk(opt_in_smem(kinv_grad_fold2_kernel<false, true>, smem));
    chk(opt_in_smem(kinv_grad_fold2_kernel<true, true>, smem));
  } else {
    chk(opt_in_smem(kinv_grad_kernel<false>, KinvG::SMEM_BYTES));
    chk(opt_in_smem(kinv_grad_kernel<true>, KinvG::SMEM_BYTES));
  }
  chk(opt_in_smem(predict_var_kernel, PredG::SMEM_BYTES));
  chk(opt_in_smem(predict_v_kernel, PredG::SMEM_BYTES));
  chk(opt_in_smem(ttv_kernel, TtvG::SMEM_BYTES));
  if (predict_grad_smem_bytes(kd) > 48 * 1024) chk(opt_in_smem(predict_grad_kernel, predict_grad_smem_bytes(kd)));
  if (e != cudaSuccess) return fail_cuda("shared-memory opt-in", e);
  int sms = 0, per_sm = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, gp->device);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, factor_kernel<false>, FAC_THREADS, FAC_SMEM_BYTES);
  if (e != cudaSuccess || per_sm < 1) return fail_cuda("factor occupancy", e);
  gp->fac_resident = sms * per_sm;
  gp->sm_count = sms;
  int per_sm_f = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, factor_kernel<true>, FAC_THREADS, FAC_SMEM_BYTES_FUSED);
  if (e != cudaSuccess || per_sm_f < 1) return fail_cuda("factor occupancy (fused panel)", e);
  gp->fac_resident_fused = sms * per_sm_f;
  if (const char* env = getenv("AVN_FAC_CTAS_PER_SM")) gp->fac_resident = sms * atoi(env);   // development knob
  gp->dev_ready = true;
  return 0;
}

// conversions + scaled inputs for B samples
// fork_y (out): set when the output column was launched on the handle's side stream -- the caller joins it with
// join_warp_y before anything reads z / dz / wstat (the factor kernel).  Only asked for by loglik_group, only taken for few
// samples of a model with a learnable output warp: that column (log, sinh-arcsinh, meanstd with block reductions: 65 us
// for one sample of c2) is the longest of the kernel and nothing before the factorisation needs it, so it runs beside the
// input columns (40 us), scale_kernel and the covariance build instead of in front of them.  Works inside a stream
// capture as well (the side stream joins the capture through the fork event).
static int run_warp(avn_gp* gp, const double* theta, int64_t B, const WsPtrs& W, int64_t npad, cudaStream_t st,
                    bool* fork_y = nullptr) {
  Phase ph(gp, AVN_PH_WARP, st);
  // columns with a warp program are staged in shared memory when they fit (see warp_kernel)
  int max_np = -1;
  for (int m = 0; m < gp->kd.d; m++)
    if (gp->progs.xw[m].nstages > 0 && gp->progs.xw[m].nparams > max_np) max_np = gp->progs.xw[m].nparams;
  if (gp->progs.yw.nstages > 0 && gp->progs.yw.nparams > max_np) max_np = gp->progs.yw.nparams;
  int64_t stage_doubles = max_np >= 0 ? gp->N * (1 + max_np) : 0;
  if (stage_doubles * 8 > kWarpStageMaxBytes) stage_doubles = 0;
  const size_t smem = (size_t)stage_doubles * 8;
  bool fork = fork_y && !gp->profiling && gp->progs.yw.nstages > 0 && B * (gp->kd.d + 1) <= gp->sm_count;
  if (fork && !gp->ystream) {
    cudaError_t e = cudaStreamCreateWithFlags(&gp->ystream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&gp->yev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&gp->yev_join, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail_cuda("warp side stream", e);
  }
  if (fork) {
    cudaError_t e = cudaEventRecord(gp->yev_fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(gp->ystream, gp->yev_fork, 0);
    if (e != cudaSuccess) return fail_cuda("warp fork", e);
    warp_kernel<<<dim3(1, (unsigned)B), WARP_THREADS, smem, gp->ystream>>>(gp->kd, gp->progs, gp->X, gp->y, (int)gp->N, (int)npad,
                                                                           theta, W, (int)stage_doubles, gp->kd.d);
    LAUNCH_CHECK("warp_kernel (output column)");
    e = cudaEventRecord(gp->yev_join, gp->ystream);
    if (e != cudaSuccess) return fail_cuda("warp join event", e);
    *fork_y = true;
  }
  warp_kernel<<<dim3((unsigned)gp->kd.d + (fork ? 0 : 1), (unsigned)B), WARP_THREADS, smem, st>>>(
      gp->kd, gp->progs, gp->X, gp->y, (int)gp->N, (int)npad, theta, W, (int)stage_doubles, 0);
  LAUNCH_CHECK("warp_kernel");
  scale_kernel<<<dim3((unsigned)((npad + 255) / 256), (unsigned)B), 256, 0, st>>>(gp->kd, (int)npad, theta, W);
  LAUNCH_CHECK("scale_kernel");
  return 0;
}
static int join_warp_y(avn_gp* gp, cudaStream_t st) {
  cudaError_t e = cudaStreamWaitEvent(st, gp->yev_join, 0);
  if (e != cudaSuccess) return fail_cuda("warp join", e);
  return 0;
}

static int run_cov(avn_gp* gp, const double* theta, int64_t B, const WsPtrs& W, int64_t npad, double* Kout,
                   cudaStream_t st) {
  Phase ph(gp, AVN_PH_COV, st);
  const KernDesc& kd = gp->kd;
  const int64_t nb = npad / TILE, ntiles = nb * (nb + 1) / 2;
  const size_t smem = cov_smem_bytes(kd);   // opted in by ensure_ready when above 48 KB
  const dim3 grid((unsigned)ntiles, (unsigned)B);
  if (kd.nkern == 1 && smem <= 48 * 1024) {
    // single-kernel models: instantiation per kernel kind (no switch / fold inside the element loop)
    switch (kd.kern[0]) {
      case AVN_RBF: cov1_kernel<AVN_RBF><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
      case AVN_MATERN52: cov1_kernel<AVN_MATERN52><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
      case AVN_MATERN32: cov1_kernel<AVN_MATERN32><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
      case AVN_EXPONENTIAL: cov1_kernel<AVN_EXPONENTIAL><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
      default: cov1_kernel<AVN_RATQUAD><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
    }
    LAUNCH_CHECK("cov1_kernel");
    return 0;
  }
  if (kd.nkern == 2 && smem <= 48 * 1024) {
    // two-kernel folds: instantiation per pair of kinds (at most one RatQuad per model: avn_gp_create)
#define AVN_COV2(A, B_) \
  case (A) * 8 + (B_): cov2_kernel<A, B_><<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout); break;
#define AVN_COV2_ROW(A) AVN_COV2(A, AVN_RBF) AVN_COV2(A, AVN_MATERN52) AVN_COV2(A, AVN_MATERN32) AVN_COV2(A, AVN_EXPONENTIAL)
    bool done = true;
    switch (kd.kern[0] * 8 + kd.kern[1]) {
      AVN_COV2_ROW(AVN_RBF) AVN_COV2(AVN_RBF, AVN_RATQUAD)
      AVN_COV2_ROW(AVN_MATERN52) AVN_COV2(AVN_MATERN52, AVN_RATQUAD)
      AVN_COV2_ROW(AVN_MATERN32) AVN_COV2(AVN_MATERN32, AVN_RATQUAD)
      AVN_COV2_ROW(AVN_EXPONENTIAL) AVN_COV2(AVN_EXPONENTIAL, AVN_RATQUAD)
      AVN_COV2_ROW(AVN_RATQUAD)
      default: done = false;
    }
#undef AVN_COV2_ROW
#undef AVN_COV2
    if (done) {
      LAUNCH_CHECK("cov2_kernel");
      return 0;
    }
  }
  cov_kernel<<<grid, 256, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.xs, W.x2, Kout);
  LAUNCH_CHECK("cov_kernel");
  return 0;
}

// Cholesky K -> L in place (kl) and T = L^-1 (t): one persistent dataflow launch (factor.cuh).
// The caller has zeroed W.lflag / W.tflag / W.ctl on this stream.
static int run_factor(avn_gp* gp, int64_t B, const WsPtrs& W, int32_t* info, int64_t npad, bool want_inverse,
                      cudaStream_t st) {
  const int nb = (int)(npad / TILE);
  const int64_t total = B * nb * nb;
  if (total > 0x7fffffffLL) return fail("run_factor: B * (N/64)^2 exceeds the task counter");
  // Few samples: fewer tile tasks per block step (B * nb) than CTAs that fit with the fused-panel footprint -- the run
  // time is the chain of diagonal tasks, which the fused panel shortens (factor.cuh); more samples: throughput mode.
  bool fused = B * nb <= gp->fac_resident_fused;
  if (const char* env = getenv("AVN_FAC_FUSE")) fused = atoi(env) != 0;   // development knob
  const int resident = fused ? gp->fac_resident_fused : gp->fac_resident;
  const int grid = (int)(total < resident ? total : resident);
  const size_t smem = fused ? FAC_SMEM_BYTES_FUSED : FAC_SMEM_BYTES;
  FactorArgs fa;
  fa.L = W.kl; fa.T = W.t; fa.fpart = W.fpart; fa.info = info; fa.z = W.z; fa.beta = W.beta;
  fa.lflag = W.lflag; fa.tflag = W.tflag; fa.sflag = W.sflag; fa.dflag = W.dflag; fa.ctl = W.ctl;
  fa.npad = (int)npad; fa.nb = nb; fa.B = (int)B; fa.n = (int)gp->N; fa.want_inverse = want_inverse ? 1 : 0;
  fa.max_spins = gp->max_spins;
  fa.fault = gp->fault;
  fa.fuse_panel = fused ? 1 : 0;
  fa.dgap = 0;   // D(.,s+1) right behind P(.,s,s+1): its first s slabs are final already, only the last one waits
  fa.prof = nullptr;
#ifdef AVN_FACTOR_PROF
  static long long* prof_dev = nullptr;
  const size_t prof_bytes = (16 + 4 * 4000) * 8;
  if (!prof_dev) cudaMalloc(&prof_dev, prof_bytes);
  cudaMemsetAsync(prof_dev, 0, prof_bytes, st);
  fa.prof = prof_dev;
#endif
  {
    Phase ph(gp, AVN_PH_FACTOR, st);
    if (fused) factor_kernel<true><<<grid, FAC_THREADS, smem, st>>>(fa);
    else factor_kernel<false><<<grid, FAC_THREADS, smem, st>>>(fa);
    LAUNCH_CHECK("factor_kernel");
  }
#ifdef AVN_FACTOR_PROF
  {
    long long h[16];
    cudaMemcpyAsync(h, prof_dev, 128, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    double tot = 0;
    for (int q = 0; q < 6; q++) tot += (double)h[q];
    fprintf(stderr, "[factor prof] grid %d  ticket %.1f%%  wait %.1f%%  gemm %.1f%%  wait_tkk %.1f%%  epilogue %.1f%%  diag %.1f%%  (cycles/CTA %.0f)\n",
            grid, 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, 100 * h[5] / tot, tot / grid);
    const double ntask_pr = (double)B * nb * (nb - 1) / 2 * (want_inverse ? 2 : 1);
    fprintf(stderr, "[factor prof] per task: diag %.0f cycles (x %d D tasks per sample), P/R epilogue %.0f cycles, P/R gemm %.0f cycles\n",
            (double)h[5] / ((double)B * nb), nb, (double)h[4] / ntask_pr, (double)h[2] / ntask_pr);
    fprintf(stderr, "[factor prof] diag fn (thread 0 cycles per D task): warp cholesky %.0f, store %.0f, trsm %.0f, syrk %.0f, sub-block inverses %.0f, T off-diagonal %.0f\n",
            (double)h[8] / ((double)B * nb), (double)h[9] / ((double)B * nb), (double)h[10] / ((double)B * nb),
            (double)h[11] / ((double)B * nb), (double)h[13] / ((double)B * nb), (double)h[12] / ((double)B * nb));
    if (B == 1 && getenv("AVN_FACTOR_TIMELINE")) {
      static long long ev[4 * 4000];
      long long cnt = h[15] < 4000 ? h[15] : 4000;
      cudaMemcpy(ev, prof_dev + 16, sizeof(long long) * 4 * cnt, cudaMemcpyDeviceToHost);
      long long t0 = 0;
      for (long long q = 0; q < cnt; q++)
        if (ev[4 * q] == 0 && ev[4 * q + 1] == 8 && ev[4 * q + 2] == 0) t0 = ev[4 * q + 3];
      const char* names[2][7] = {{"D start", "D update done", "D chol+inv done", "D published", "D pipeline done", "D flag seen", "D tile staged"},
                                 {"P start", "P gemm done", "P has T_kk", "P published", "", "", ""}};
      for (int kq = 8; kq <= 10; kq++)
        for (int ty = 0; ty < 2; ty++)
          for (int evn = 0; evn < 7; evn++)
            for (long long q = 0; q < cnt; q++)
              if (ev[4 * q] == ty && ev[4 * q + 1] == kq && ev[4 * q + 2] == evn)
                fprintf(stderr, "[timeline] k=%d %-18s %8.2f us\n", kq, names[ty][evn], (ev[4 * q + 3] - t0) * 1e-3);
    }
    fprintf(stderr, "[factor prof] D task: stage %.0f cycles, chol+inverse %.0f cycles, store+publish %.0f cycles\n",
            (double)h[7] / ((double)B * nb), (double)h[6] / ((double)B * nb), (double)h[5] / ((double)B * nb));
  }
#endif
  return 0;
}

// beta = T z (+ partial sums of beta^T beta), alpha = T^T beta
static int run_beta_alpha(avn_gp* gp, int64_t B, const WsPtrs& W, int64_t npad, bool want_alpha, cudaStream_t st) {
  const dim3 grid((unsigned)(npad / TILE), (unsigned)B);
  {
    Phase ph(gp, AVN_PH_BETA, st);
    beta_reduce_kernel<<<grid, 64, 0, st>>>(W.kl, (int)npad, W.beta, W.fpart);
    LAUNCH_CHECK("beta_reduce_kernel");
  }
  if (want_alpha) {
    Phase ph(gp, AVN_PH_ALPHA, st);
    alpha_kernel<<<grid, ALPHA_THREADS, 0, st>>>(W.t, W.beta, (int)npad, W.alpha);
    LAUNCH_CHECK("alpha_kernel");
  }
  return 0;
}

static cudaError_t zero_flags(const WsPtrs& W, int64_t B, int64_t nb, cudaStream_t st) {
  return cudaMemsetAsync(W.lflag, 0, sizeof(int32_t) * (size_t)(4 * B * nb + 64 + kKinvSchedInts), st);
}

extern "C" int avn_gp_cov(avn_gp* gp, const double* theta_dev, int64_t B, double* K_dev, void* ws_dev, size_t ws_bytes,
                          void* stream) {
  if (!gp || !theta_dev || !K_dev || !ws_dev) return fail("avn_gp_cov: null argument");
  if (gp->N < 1) return fail("avn_gp_cov: set_data first");
  if (B < 1 || B > 65535) return fail("avn_gp_cov: B out of range [1,65535]");
  avn_ws_layout L;
  layout(gp, B, &L);
  if (ws_bytes < (size_t)L.total) return fail("avn_gp_cov: workspace too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WsPtrs W = ws_ptrs(L, ws_dev, B);
  int rc = run_warp(gp, theta_dev, B, W, L.npad, st);
  if (rc) return rc;
  return run_cov(gp, theta_dev, B, W, L.npad, K_dev, st);
}

static WsPtrs ws_offset(const WsPtrs& W, const avn_gp* gp, const avn_ws_layout& L, int64_t b0) {
  const KernDesc& kd = gp->kd;
  const int64_t npad = L.npad, nb = L.nb, ntiles = nb * (nb + 1) / 2;
  WsPtrs p = W;
  p.xw += b0 * npad * kd.d;
  p.dxw += b0 * npad * kd.d * MAXWP;
  p.xs += b0 * kd.nkern * npad * kd.d;
  p.x2 += b0 * kd.nkern * npad;
  p.z += b0 * npad;
  p.dz += b0 * npad * MAXWP;
  p.wstat += b0 * WSTAT;
  p.kl += b0 * npad * npad;
  p.t += b0 * npad * npad;
  p.beta += b0 * npad;
  p.alpha += b0 * npad;
  p.gpart += b0 * ntiles * MAXACC;
  p.gxpart += b0 * nb * npad * kd.d;
  p.fpart += b0 * nb * 2;
  p.lflag += b0 * nb;
  p.tflag += b0 * nb;
  p.sflag += b0 * nb;
  p.dflag += b0 * nb;
  return p;
}

// optin_only: the shared-memory opt-in of this instantiation (ensure_ready), no launch.  single > 0: the single-sample
// kernel on `single` CTAs (as many as are resident at once), tiles dealt by placement.
template <int KIND, bool GX>
static int launch_kinv_fast_t(bool optin_only, dim3 grid, size_t smem, cudaStream_t st, const KernDesc& kd, int N, int npad,
                              const double* theta, const WsPtrs& W, int single, int nsm) {
  if (optin_only) {
    cudaError_t e = opt_in_smem(kinv_grad_fast_kernel<KIND, GX>, smem);
    if (e == cudaSuccess) e = opt_in_smem(kinv_grad_fast_single_kernel<KIND, GX>, smem);
    if (e != cudaSuccess) return fail_cuda("kinv_grad_fast smem opt-in", e);
    return 0;
  }
  if (single > 0) {
    const int ntiles = (int)grid.x;
    kinv_grad_fast_single_kernel<KIND, GX><<<(unsigned)(ntiles < single ? ntiles : single), KinvG2::NTHREADS, smem, st>>>(
        kd, N, npad, theta, W.t, W.alpha, W.xw, W.xs, W.x2, W.gpart, W.gxpart, ntiles, nsm, W.ksched);
    return 0;
  }
  kinv_grad_fast_kernel<KIND, GX><<<grid, KinvG2::NTHREADS, smem, st>>>(kd, N, npad, theta, W.t, W.alpha, W.xw, W.xs, W.x2,
                                                                        W.gpart, W.gxpart);
  return 0;
}

static int launch_kinv_fast(bool optin_only, int kind, bool gx, dim3 grid, size_t smem, cudaStream_t st, const KernDesc& kd,
                            int N, int npad, const double* theta, const WsPtrs& W, int single, int nsm) {
#define AVN_DISPATCH(K)                                                                                        \
  case K:                                                                                                      \
    return gx ? launch_kinv_fast_t<K, true>(optin_only, grid, smem, st, kd, N, npad, theta, W, single, nsm)    \
              : launch_kinv_fast_t<K, false>(optin_only, grid, smem, st, kd, N, npad, theta, W, single, nsm);
  switch (kind) {
    AVN_DISPATCH(AVN_RBF)
    AVN_DISPATCH(AVN_MATERN52)
    AVN_DISPATCH(AVN_MATERN32)
    AVN_DISPATCH(AVN_EXPONENTIAL)
    AVN_DISPATCH(AVN_RATQUAD)
  }
#undef AVN_DISPATCH
  return fail("launch_kinv_fast: unknown kernel kind");
}

// the whole evaluation for samples [b0, b0+Bg) on one stream
static int loglik_group(avn_gp* gp, const double* theta, int64_t Bg, double* ll, double* grad, int32_t* info,
                        const WsPtrs& W, const avn_ws_layout& L, cudaStream_t st, bool single_sample = false,
                        bool single_group = false) {
  const KernDesc& kd = gp->kd;
  const int64_t npad = L.npad, nb = L.nb, ntiles = nb * (nb + 1) / 2;
  const bool want_grad = grad != nullptr;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int32_t) * Bg, st);
  if (e != cudaSuccess) return fail_cuda("memset info", e);
  bool fork_y = false;
  int rc = run_warp(gp, theta, Bg, W, npad, st, single_group ? &fork_y : nullptr);
  if (rc) {
    if (fork_y) join_warp_y(gp, st);   // a forked stream must rejoin (a capture could not end otherwise)
    return rc;
  }
  rc = run_cov(gp, theta, Bg, W, npad, W.kl, st);
  if (fork_y) {
    const int rj = join_warp_y(gp, st);
    if (rc == 0) rc = rj;
  }
  if (rc) return rc;
  rc = run_factor(gp, Bg, W, info, npad, true, st);   // beta = T z needs the inverse also without a gradient
  if (rc) return rc;
  rc = run_beta_alpha(gp, Bg, W, npad, want_grad, st);
  if (rc) return rc;
  if (want_grad) {
    Phase ph(gp, AVN_PH_KINV_GRAD, st);
    const dim3 grid((unsigned)ntiles, (unsigned)Bg);
    if (kd.nkern == 1) {
      // single-kernel model: DMMA epilogue specialised on the kernel kind
      // a single sample (the call's B, not a group's): tiles dealt by placement over the resident CTAs (kinv_fast.cuh)
      rc = launch_kinv_fast(false, kd.kern[0], gp->has_xwarp, grid, kinv_fast_smem_bytes(kd), st, kd, (int)gp->N, (int)npad,
                            theta, W, single_sample ? 3 * (gp->sm_count < 256 ? gp->sm_count : 256) : 0,
                            gp->sm_count < 256 ? gp->sm_count : 256);   // (the scheduler block holds 256 SM counters, 768 claim words)
      if (rc) return rc;
    } else if (kd.nkern == 2 && !(kd.kern[0] == AVN_RATQUAD && kd.kern[1] == AVN_RATQUAD)) {
      // two-kernel sum / product: DMMA epilogue with the fold's product rule (kinv_fold.cuh)
      const size_t smem = kinv_fold_smem_bytes(kd);
      const bool rq = kd.kern[0] == AVN_RATQUAD || kd.kern[1] == AVN_RATQUAD;
#define AVN_FOLD2(GX, RQ)                                                                                              \
  kinv_grad_fold2_kernel<GX, RQ><<<grid, KinvG2::NTHREADS, smem, st>>>(kd, (int)gp->N, (int)npad, theta, W.t, W.alpha, \
                                                                       W.xw, W.xs, W.x2, W.gpart, W.gxpart)
      if (gp->has_xwarp) {
        if (rq) AVN_FOLD2(true, true); else AVN_FOLD2(true, false);
      } else {
        if (rq) AVN_FOLD2(false, true); else AVN_FOLD2(false, false);
      }
#undef AVN_FOLD2
    } else if (gp->has_xwarp) {
      kinv_grad_kernel<true><<<grid, KinvG::NTHREADS, KinvG::SMEM_BYTES, st>>>(
          kd, (int)gp->N, (int)npad, theta, W.t, W.alpha, W.xw, W.gpart, W.gxpart);
    } else {
      kinv_grad_kernel<false><<<grid, KinvG::NTHREADS, KinvG::SMEM_BYTES, st>>>(
          kd, (int)gp->N, (int)npad, theta, W.t, W.alpha, W.xw, W.gpart, W.gxpart);
    }
    LAUNCH_CHECK("kinv_grad_kernel");
  }
  Phase phf(gp, AVN_PH_FINALIZE, st);
  if (want_grad && kd.n_iw > 0) {
    gx_reduce_kernel<<<dim3((unsigned)((npad * kd.d + 255) / 256), (unsigned)Bg), 256, 0, st>>>((int)npad, kd.d, W.gxpart);
    LAUNCH_CHECK("gx_reduce_kernel");
  }
  finalize_kernel<<<dim3((unsigned)Bg, (unsigned)finalize_grid_y(kd.n_iw, kd.n_cw, want_grad ? 1 : 0)), 256, 0, st>>>(
      kd, gp->progs, (int)gp->N, (int)npad, (int)ntiles, want_grad ? 1 : 0,
                                                theta, W, info, ll, grad);
  LAUNCH_CHECK("finalize_kernel");
  return 0;
}

extern "C" int avn_gp_set_streams(avn_gp* gp, int max_groups) {
  if (!gp) return fail("avn_gp_set_streams: null handle");
  if (max_groups < 1 || max_groups > 8) return fail("avn_gp_set_streams: max_groups out of range [1,8]");
  gp->max_groups = max_groups;
  return 0;
}

extern "C" int avn_gp_loglik_grad(avn_gp* gp, const double* theta_dev, int64_t B, double* ll_dev, double* grad_dev,
                                  int32_t* info_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!gp || !theta_dev || !ll_dev || !info_dev || !ws_dev) return fail("avn_gp_loglik_grad: null argument");
  if (gp->N < 1) return fail("avn_gp_loglik_grad: set_data first");
  if (B < 1 || B > 65535) return fail("avn_gp_loglik_grad: B out of range [1,65535]");
  avn_ws_layout L;
  layout(gp, B, &L);
  if (ws_bytes < (size_t)L.total) return fail("avn_gp_loglik_grad: workspace too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WsPtrs W = ws_ptrs(L, ws_dev, B);
  const int P = gp->kd.P;
  // Samples are independent: split them into groups that run the whole sequence on internal streams, so the
  // latency-bound steps of one group (diagonal blocks, short panels, tails) overlap the DMMA-bound steps of
  // another.  Phase timing needs one ordered stream, so profiling forces a single group.
  int G = gp->profiling ? 1 : gp->max_groups;
  if (B < 8 * G) G = (int)(B / 8 > 0 ? B / 8 : 1);
  {
    cudaError_t e = zero_flags(W, B, L.nb, st);
    if (e != cudaSuccess) return fail_cuda("memset flags", e);
  }
  if (G <= 1) return loglik_group(gp, theta_dev, B, ll_dev, grad_dev, info_dev, W, L, st, B == 1, true);
  for (int g = 0; g < G; g++)
    if (!gp->gstream[g]) {
      cudaError_t e = cudaStreamCreateWithFlags(&gp->gstream[g], cudaStreamNonBlocking);
      if (e != cudaSuccess) return fail_cuda("cudaStreamCreate", e);
    }
  for (int g = 0; g <= G; g++)
    if (!gp->gev[g]) {
      cudaError_t e = cudaEventCreateWithFlags(&gp->gev[g], cudaEventDisableTiming);
      if (e != cudaSuccess) return fail_cuda("cudaEventCreate", e);
    }
  cudaError_t e = cudaEventRecord(gp->gev[0], st);
  if (e != cudaSuccess) return fail_cuda("fork event", e);
  int rc = 0;
  for (int g = 0; g < G; g++) {
    const int64_t b0 = B * g / G, b1 = B * (g + 1) / G;
    cudaStream_t gs = gp->gstream[g];
    cudaStreamWaitEvent(gs, gp->gev[0], 0);
    WsPtrs Wg = ws_offset(W, gp, L, b0);
    Wg.ctl = W.ctl + 8 * g;
    rc = loglik_group(gp, theta_dev + b0 * P, b1 - b0, ll_dev + b0, grad_dev ? grad_dev + b0 * P : nullptr,
                      info_dev + b0, Wg, L, gs);
    cudaEventRecord(gp->gev[1 + g], gs);
    cudaStreamWaitEvent(st, gp->gev[1 + g], 0);  // join even on error so the caller's stream stays ordered
    if (rc) break;
  }
  return rc;
}

// ---- host-buffer evaluation (what find_MAP / pm.sample call once per step: NumPy point in, logp + dlogp out) ----------
extern "C" size_t avn_gp_host_staging_bytes(const avn_gp* gp, int64_t B) {
  if (!gp || B < 1) return 0;
  const int64_t P = gp->kd.P;
  return (size_t)(align_up(B * P * 8, 256) + align_up(B * (P + 2) * 8, 256));
}

extern "C" int avn_gp_host_wait(avn_gp* gp) {
  if (!gp) return fail("avn_gp_host_wait: null handle");
  if (!gp->hstream) return 0;
  DevGuard guard(gp);
  cudaError_t e = cudaStreamSynchronize(gp->hstream);
  if (e != cudaSuccess) return fail_cuda("avn_gp_host_wait", e);
  return 0;
}

extern "C" int avn_gp_loglik_grad_host(avn_gp* gp, const double* theta_host, int64_t B, double* out_host, int32_t flags,
                                       void* staging_dev, size_t staging_bytes, void* ws_dev, size_t ws_bytes, void* stream) {
  const int want_grad = flags & 1;
  const bool no_wait = (flags & 2) != 0;
  if (!gp || !theta_host || !out_host || !staging_dev || !ws_dev) return fail("avn_gp_loglik_grad_host: null argument");
  if (gp->N < 1) return fail("avn_gp_loglik_grad_host: set_data first");
  if (B < 1 || B > 65535) return fail("avn_gp_loglik_grad_host: B out of range [1,65535]");
  if (staging_bytes < avn_gp_host_staging_bytes(gp, B)) return fail("avn_gp_loglik_grad_host: staging buffer too small");
  avn_ws_layout L;
  layout(gp, B, &L);
  if (ws_bytes < (size_t)L.total) return fail("avn_gp_loglik_grad_host: workspace too small");
  ENTER_DEVICE(gp);
  const int64_t P = gp->kd.P;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (!gp->hstream) {
    e = cudaStreamCreateWithFlags(&gp->hstream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&gp->hev, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail_cuda("host-call stream", e);
  }
  double* theta_dev = static_cast<double*>(staging_dev);
  double* packed = reinterpret_cast<double*>(static_cast<char*>(staging_dev) + align_up(B * P * 8, 256));
  double* ll_dev = packed;
  double* grad_dev = packed + B;
  int32_t* info_dev = reinterpret_cast<int32_t*>(packed + B + B * P);
  const avn_gp::HostKey key = {theta_host, out_host, staging_dev, ws_dev, gp->X, gp->y, B, gp->N, want_grad ? 1 : 0,
                               gp->max_spins, gp->fault};
  if (!gp->hexec || !(key == gp->hkey)) {
    // capture: H2D of the points, flags, the 8-9 launches of the evaluation, D2H of the packed results
    // (a replay the caller has not waited for yet finishes first: its graph is about to be destroyed)
    e = cudaStreamSynchronize(gp->hstream);
    if (e != cudaSuccess) return fail_cuda("host-call stream", e);
    host_graph_drop(gp);
    const bool prof = gp->profiling;
    gp->profiling = false;   // phase events cannot be recorded into a capture
    e = cudaStreamBeginCapture(gp->hstream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { gp->profiling = prof; return fail_cuda("begin capture", e); }
    cudaStream_t hs = gp->hstream;
    gp->launches = 0;
    WsPtrs W = ws_ptrs(L, ws_dev, B);
    int rc = 0;
    e = cudaMemcpyAsync(theta_dev, theta_host, (size_t)(B * P * 8), cudaMemcpyHostToDevice, hs);
    if (e == cudaSuccess) e = cudaMemsetAsync(info_dev, 0, (size_t)(B * 8), hs);   // whole [B] doubles slot of the int32 info
    if (e == cudaSuccess) e = zero_flags(W, B, L.nb, hs);
    if (e == cudaSuccess) rc = loglik_group(gp, theta_dev, B, ll_dev, want_grad ? grad_dev : nullptr, info_dev, W, L, hs, B == 1, true);
    if (e == cudaSuccess && rc == 0)
      e = cudaMemcpyAsync(out_host, packed, (size_t)(B * (P + 2) * 8), cudaMemcpyDeviceToHost, hs);
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(hs, &graph);
    gp->profiling = prof;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      return fail_cuda("capture", e != cudaSuccess ? e : e2);
    }
    e = cudaGraphInstantiate(&gp->hexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { gp->hexec = nullptr; return fail_cuda("graph instantiate", e); }
    gp->hkey = key;
    gp->hlaunches = gp->launches;
  }
  gp->launches = gp->hlaunches;
  // replay behind whatever the caller has enqueued on its stream, wait for the results
  e = cudaEventRecord(gp->hev, st);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(gp->hstream, gp->hev, 0);
  if (e == cudaSuccess) e = cudaGraphLaunch(gp->hexec, gp->hstream);
  if (e == cudaSuccess && !no_wait) e = cudaStreamSynchronize(gp->hstream);
  if (e != cudaSuccess) return fail_cuda("graph launch", e);
  return 0;
}

// ---- predict -------------------------------------------------------------------------------------
struct StateLayout {
  int64_t hyp, alpha, xs, x2, t, total;  // byte offsets
};
static StateLayout state_layout(const avn_gp* gp) {
  const KernDesc& kd = gp->kd;
  const int64_t npad = npad_of(gp->N);
  StateLayout S;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  S.hyp = take(sizeof(HypS));
  S.alpha = take(npad * 8);
  S.xs = take((int64_t)kd.nkern * npad * kd.d * 8);
  S.x2 = take((int64_t)kd.nkern * npad * 8);
  S.t = take(npad * npad * 8);
  S.total = off;
  return S;
}

extern "C" size_t avn_gp_state_bytes(const avn_gp* gp) {
  if (!gp || gp->N < 1) return 0;
  return (size_t)state_layout(gp).total;
}

// also the abort check of avn_gp_factorize: a timed-out wait of the factor kernel (ctl[1] != 0) becomes info = -1
__global__ void hyp_store_kernel(KernDesc kd, const double* __restrict__ theta, HypS* __restrict__ out,
                                 const int32_t* __restrict__ ctl, int32_t* __restrict__ info) {
  __shared__ HypS h;
  if (threadIdx.x == 0 && ctl[1] != 0) info[0] = -1;
  for (int e = threadIdx.x; e < (int)(sizeof(HypS) / 8); e += blockDim.x) reinterpret_cast<double*>(&h)[e] = 0.0;
  __syncthreads();
  load_hyp(h, kd, theta);
  __syncthreads();
  for (int e = threadIdx.x; e < (int)(sizeof(HypS) / 8); e += blockDim.x)
    reinterpret_cast<double*>(out)[e] = reinterpret_cast<double*>(&h)[e];
}

extern "C" int avn_gp_factorize(avn_gp* gp, const double* theta_dev, void* state_dev, size_t state_bytes,
                                int32_t* info_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!gp || !theta_dev || !state_dev || !info_dev || !ws_dev) return fail("avn_gp_factorize: null argument");
  if (gp->N < 1) return fail("avn_gp_factorize: set_data first");
  if (gp->has_xwarp || gp->kd.n_cw > 0)
    return fail("avn_gp_factorize: predict works on converted data; bake the warps on the host first (gpmcmc.py:364-399)");
  avn_ws_layout L;
  const int64_t B = 1;
  layout(gp, B, &L);
  StateLayout S = state_layout(gp);
  if (ws_bytes < (size_t)L.total) return fail("avn_gp_factorize: workspace too small");
  if (state_bytes < (size_t)S.total) return fail("avn_gp_factorize: state buffer too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WsPtrs W = ws_ptrs(L, ws_dev, B);
  char* sb = static_cast<char*>(state_dev);
  // T, alpha, xs, x2 are produced directly inside the state buffer
  W.t = reinterpret_cast<double*>(sb + S.t);
  W.alpha = reinterpret_cast<double*>(sb + S.alpha);
  W.xs = reinterpret_cast<double*>(sb + S.xs);
  W.x2 = reinterpret_cast<double*>(sb + S.x2);
  const int64_t npad = L.npad;
  cudaError_t e = cudaMemsetAsync(info_dev, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return fail_cuda("memset info", e);
  int rc = run_warp(gp, theta_dev, 1, W, npad, st);
  if (rc) return rc;
  rc = run_cov(gp, theta_dev, 1, W, npad, W.kl, st);
  if (rc) return rc;
  e = zero_flags(W, 1, L.nb, st);
  if (e != cudaSuccess) return fail_cuda("memset flags", e);
  rc = run_factor(gp, 1, W, info_dev, npad, true, st);
  if (rc) return rc;
  rc = run_beta_alpha(gp, 1, W, npad, true, st);
  if (rc) return rc;
  hyp_store_kernel<<<1, 128, 0, st>>>(gp->kd, theta_dev, reinterpret_cast<HypS*>(sb + S.hyp), W.ctl, info_dev);
  LAUNCH_CHECK("hyp_store_kernel");
  return 0;
}

// cross-covariance panel + latent mean: single-kernel models take the register-blocked instantiation of their kind
static void launch_kxs(const KernDesc& kd, dim3 grid, cudaStream_t st, int N, int npad, const HypS* hyp, const double* xs,
                       const double* x2, const double* alpha, const double* Xs_dev, int64_t M, int64_t m0, int cols,
                       double* Kxs, double* out_mean_dev, int ns, double* mu_part) {
  if (kd.nkern == 1) {
    const size_t smem = kxs1_smem_doubles(kd.d) * 8;
#define AVN_KXS1(K) \
  case K: kxs1_kernel<K><<<grid, 256, smem, st>>>(kd, N, npad, hyp, xs, x2, alpha, Xs_dev, M, m0, cols, Kxs, out_mean_dev, ns, mu_part); return;
    switch (kd.kern[0]) {
      AVN_KXS1(AVN_RBF)
      AVN_KXS1(AVN_MATERN52)
      AVN_KXS1(AVN_MATERN32)
      AVN_KXS1(AVN_EXPONENTIAL)
      AVN_KXS1(AVN_RATQUAD)
    }
#undef AVN_KXS1
  }
  kxs_kernel<<<grid, 256, kxs_smem_bytes(kd), st>>>(kd, N, npad, hyp, xs, x2, alpha, Xs_dev, M, m0, cols, Kxs, out_mean_dev,
                                                     ns, mu_part);
}

static const int64_t kPanelCols = 148 * 64 * AVN_PRED_CTAS;  // test points per K_xs panel: one wave of the predict kernels

// Small test batches (BO candidates, refine / inverse-problem starts): fewer 64-point column blocks than CTA slots
// (2 resident CTAs x 148 SMs).  The training rows are then split over blockIdx.y as well -- as many splits as still fit
// in ONE wave (a second, partly filled wave costs more than it gains: measured) -- and a finish kernel sums the partials
// in fixed order.  A function of the TOTAL M, so that the result does not depend on how a call is cut into panels.
static int row_split(int64_t M, int64_t npad) {
  const int64_t nblk = (M + TILE - 1) / TILE, nb = npad / TILE, slots = AVN_PRED_CTAS * 148;
  int64_t ns = slots / nblk;
  if (ns > nb) ns = nb;
  return (int)(ns < 1 ? 1 : ns);
}

extern "C" size_t avn_gp_predict_workspace_bytes(const avn_gp* gp, int64_t M) {
  if (!gp || gp->N < 1 || M < 1) return 0;
  const int64_t npad = npad_of(gp->N);
  int64_t cols = align_up(M < kPanelCols ? M : kPanelCols, TILE);
  const int64_t ns = row_split(M, npad);
  return (size_t)((npad + (ns > 1 ? 2 * ns : 0)) * cols * 8);
}

extern "C" int avn_gp_predict(avn_gp* gp, const void* state_dev, const double* Xs_dev, int64_t M,
                              const avn_epilogue* epi, const double* mean_add_dev, double* out_mean_dev,
                              double* out_var_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!gp || !state_dev || !Xs_dev || !epi || !out_mean_dev || !out_var_dev || !ws_dev)
    return fail("avn_gp_predict: null argument");
  if (gp->N < 1) return fail("avn_gp_predict: set_data first");
  if (M < 1) return fail("avn_gp_predict: M must be >= 1");
  if (epi->mode < 0 || epi->mode > 2) return fail("avn_gp_predict: bad epilogue mode");
  if (epi->mode != 0 && (epi->deg < 1 || epi->deg > AVN_MAX_GH)) return fail("avn_gp_predict: deg out of range [1,32]");
  const KernDesc& kd = gp->kd;
  const int64_t npad = npad_of(gp->N);
  StateLayout S = state_layout(gp);
  const int ns = row_split(M, npad);
  const int64_t cols_cap = (int64_t)(ws_bytes / ((npad + (ns > 1 ? 2 * ns : 0)) * 8)) / TILE * TILE;
  if (cols_cap < TILE) return fail("avn_gp_predict: workspace too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* sb = static_cast<const char*>(state_dev);
  const HypS* hyp = reinterpret_cast<const HypS*>(sb + S.hyp);
  const double* alpha = reinterpret_cast<const double*>(sb + S.alpha);
  const double* xs = reinterpret_cast<const double*>(sb + S.xs);
  const double* x2 = reinterpret_cast<const double*>(sb + S.x2);
  const double* T = reinterpret_cast<const double*>(sb + S.t);
  double* Kxs = static_cast<double*>(ws_dev);
  for (int64_t m0 = 0; m0 < M; m0 += cols_cap) {
    const int64_t cols = align_up((M - m0) < cols_cap ? (M - m0) : cols_cap, TILE);
    const unsigned nblk = (unsigned)(cols / TILE);
    double* mu_part = Kxs + npad * cols;           // [ns][cols]   (row-split runs only)
    double* vpart = mu_part + (int64_t)ns * cols;  // [ns][cols]
    {
      Phase ph(gp, AVN_PH_KXS, st);
      launch_kxs(kd, dim3(nblk, ns), st, (int)gp->N, (int)npad, hyp, xs, x2, alpha, Xs_dev, M, m0, (int)cols, Kxs,
                 out_mean_dev, ns, mu_part);
      LAUNCH_CHECK("kxs_kernel");
    }
    Phase ph2(gp, AVN_PH_PREDICT_VAR, st);
    predict_var_kernel<<<dim3(nblk, ns), PredG::NTHREADS, PredG::SMEM_BYTES, st>>>(
        kd, (int)npad, hyp, T, Kxs, (int)cols, M, m0, *epi, mean_add_dev, out_mean_dev, out_var_dev, ns, vpart);
    LAUNCH_CHECK("predict_var_kernel");
    if (ns > 1) {
      predict_finish_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(kd, hyp, ns, mu_part, vpart, (int)cols, M, m0, *epi,
                                                                           1, mean_add_dev, out_mean_dev, out_var_dev);
      LAUNCH_CHECK("predict_finish_kernel");
    }
  }
  return 0;
}

// ---- predict with gradients w.r.t. the query points (BO refine) -------------------------------------
extern "C" size_t avn_gp_predict_grad_workspace_bytes(const avn_gp* gp, int64_t M) {
  if (!gp || gp->N < 1 || M < 1) return 0;
  const int64_t npad = npad_of(gp->N);
  int64_t cols = align_up(M < kPanelCols ? M : kPanelCols, TILE);
  const int64_t ns = row_split(M, npad);
  return (size_t)((2 * npad + (ns > 1 ? ns * (2 + 2 * gp->kd.d) : 0)) * cols * 8);
}

extern "C" int avn_gp_predict_grad(avn_gp* gp, const void* state_dev, const double* Xs_dev, int64_t M,
                                   const avn_epilogue* epi, int32_t pred_noise, const double* mean_add_dev,
                                   const double* dmean_add_dev, double* out_mean_dev, double* out_var_dev,
                                   double* out_dmean_dev, double* out_dvar_dev, void* ws_dev, size_t ws_bytes,
                                   void* stream) {
  if (!gp || !state_dev || !Xs_dev || !epi || !out_mean_dev || !out_var_dev || !out_dmean_dev || !out_dvar_dev || !ws_dev)
    return fail("avn_gp_predict_grad: null argument");
  if (gp->N < 1) return fail("avn_gp_predict_grad: set_data first");
  if (M < 1) return fail("avn_gp_predict_grad: M must be >= 1");
  if (epi->mode < 0 || epi->mode > 2) return fail("avn_gp_predict_grad: bad epilogue mode");
  if (epi->mode != 0 && (epi->deg < 1 || epi->deg > AVN_MAX_GH))
    return fail("avn_gp_predict_grad: deg out of range [1,32]");
  const KernDesc& kd = gp->kd;
  const int64_t npad = npad_of(gp->N);
  StateLayout S = state_layout(gp);
  const int ns = row_split(M, npad);
  const int64_t per_col = 2 * npad + (ns > 1 ? (int64_t)ns * (2 + 2 * kd.d) : 0);
  const int64_t cols_cap = (int64_t)(ws_bytes / (per_col * 8)) / TILE * TILE;
  if (cols_cap < TILE) return fail("avn_gp_predict_grad: workspace too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* sb = static_cast<const char*>(state_dev);
  const HypS* hyp = reinterpret_cast<const HypS*>(sb + S.hyp);
  const double* alpha = reinterpret_cast<const double*>(sb + S.alpha);
  const double* xs = reinterpret_cast<const double*>(sb + S.xs);
  const double* x2 = reinterpret_cast<const double*>(sb + S.x2);
  const double* T = reinterpret_cast<const double*>(sb + S.t);
  const size_t smem_pg = predict_grad_smem_bytes(kd);   // opted in by ensure_ready when above 48 KB
  avn_epilogue latent = *epi;
  latent.mode = 0;
  for (int64_t m0 = 0; m0 < M; m0 += cols_cap) {
    const int64_t cols = align_up((M - m0) < cols_cap ? (M - m0) : cols_cap, TILE);
    const unsigned nblk = (unsigned)(cols / TILE);
    double* Kxs = static_cast<double*>(ws_dev);          // K_xs, later W = K^-1 K_xs
    double* V = Kxs + npad * cols;
    double* mu_part = V + npad * cols;             // row-split runs only: [ns][cols]
    double* vpart = mu_part + (int64_t)ns * cols;  // [ns][cols]
    double* gpart = vpart + (int64_t)ns * cols;    // [ns][cols][2 d]
    {
      Phase ph(gp, AVN_PH_KXS, st);
      launch_kxs(kd, dim3(nblk, ns), st, (int)gp->N, (int)npad, hyp, xs, x2, alpha, Xs_dev, M, m0, (int)cols, Kxs,
                 out_mean_dev, ns, mu_part);
      LAUNCH_CHECK("kxs_kernel");
    }
    Phase ph2(gp, AVN_PH_PREDICT_VAR, st);
    predict_v_kernel<<<dim3(nblk, ns), PredG::NTHREADS, PredG::SMEM_BYTES, st>>>(
        kd, (int)npad, hyp, T, Kxs, (int)cols, M, m0, pred_noise ? 1 : 0, V, out_var_dev, ns, vpart);
    LAUNCH_CHECK("predict_v_kernel");
    if (ns > 1) {
      predict_finish_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(kd, hyp, ns, mu_part, vpart, (int)cols, M, m0,
                                                                           latent, pred_noise ? 1 : 0, nullptr,
                                                                           out_mean_dev, out_var_dev);
      LAUNCH_CHECK("predict_finish_kernel");
    }
    ttv_kernel<<<dim3(nblk, (unsigned)(npad / TILE)), TtvG::NTHREADS, TtvG::SMEM_BYTES, st>>>((int)npad, T, V, (int)cols,
                                                                                             Kxs);
    LAUNCH_CHECK("ttv_kernel");
    predict_grad_kernel<<<dim3(nblk, ns), 256, smem_pg, st>>>(kd, (int)gp->N, (int)npad, hyp, xs, x2, alpha, Kxs, (int)cols,
                                                              Xs_dev, M, m0, *epi, mean_add_dev, dmean_add_dev, out_mean_dev,
                                                              out_var_dev, out_dmean_dev, out_dvar_dev, ns, gpart);
    LAUNCH_CHECK("predict_grad_kernel");
    if (ns > 1) {
      predict_grad_finish_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(
          kd.d, ns, gpart, (int)cols, M, m0, *epi, mean_add_dev, dmean_add_dev, out_mean_dev, out_var_dev, out_dmean_dev,
          out_dvar_dev);
      LAUNCH_CHECK("predict_grad_finish_kernel");
    }
  }
  return 0;
}

// ---- rank-1 append ------------------------------------------------------------------------
extern "C" size_t avn_gp_append_workspace_bytes(const avn_gp* gp) {
  if (!gp || gp->N < 1) return 0;
  const int64_t npad = npad_of(gp->N);
  return (size_t)((3 * npad + 2 * (npad / TILE) + (npad + 255) / 256 + 32) * 8);
}

extern "C" int avn_gp_append(avn_gp* gp, void* state_dev, size_t state_bytes, const double* xnew_dev,
                             const double* znew_dev, int32_t* info_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  if (!gp || !state_dev || !xnew_dev || !znew_dev || !info_dev || !ws_dev) return fail("avn_gp_append: null argument");
  if (gp->N < 1) return fail("avn_gp_append: set_data first");
  if (gp->has_xwarp || gp->kd.n_cw > 0) return fail("avn_gp_append: works on converted data (as avn_gp_factorize)");
  const KernDesc& kd = gp->kd;
  const int64_t N = gp->N, npad = npad_of(N), nb = npad / TILE;
  if (N == npad) return 1;  // the padded slab is full: the caller refactorises with N + 1 points
  StateLayout S = state_layout(gp);
  if (state_bytes < (size_t)S.total) return fail("avn_gp_append: state buffer too small");
  if (ws_bytes < avn_gp_append_workspace_bytes(gp)) return fail("avn_gp_append: workspace too small");
  ENTER_DEVICE(gp);
  gp->launches = 0;
  phases_reset(gp);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* sb = static_cast<char*>(state_dev);
  const HypS* hyp = reinterpret_cast<const HypS*>(sb + S.hyp);
  double* alpha = reinterpret_cast<double*>(sb + S.alpha);
  double* xs = reinterpret_cast<double*>(sb + S.xs);
  double* x2 = reinterpret_cast<double*>(sb + S.x2);
  double* T = reinterpret_cast<double*>(sb + S.t);
  double* kvec = static_cast<double*>(ws_dev);
  double* v = kvec + npad;
  double* w = v + npad;
  double* fpart = w + npad;        // [nb][2], slot 0 = block partial of |v|^2
  double* mu_part = fpart + 2 * nb;  // block partials of k^T alpha
  const unsigned nblk = (unsigned)((npad + 255) / 256);
  cudaError_t e = cudaMemsetAsync(info_dev, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return fail_cuda("memset info", e);
  kvec_kernel<<<nblk, 256, 0, st>>>(kd, (int)N, (int)npad, hyp, xs, x2, alpha, xnew_dev, kvec, mu_part);
  LAUNCH_CHECK("kvec_kernel");
  beta_kernel<<<dim3((unsigned)nb, 1), 256, 0, st>>>(T, kvec, (int)npad, v, fpart);
  LAUNCH_CHECK("beta_kernel");
  alpha_kernel<<<dim3((unsigned)nb, 1), ALPHA_THREADS, 0, st>>>(T, v, (int)npad, w);
  LAUNCH_CHECK("alpha_kernel");
  append_kernel<<<nblk, 256, 0, st>>>(kd, (int)N, (int)npad, hyp, xnew_dev, znew_dev, w, mu_part, fpart, T, alpha, xs, x2,
                                      info_dev);
  LAUNCH_CHECK("append_kernel");
  return 0;
}

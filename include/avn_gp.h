/*
 * avn_gp.h -- C ABI of the B200-native Gaussian-process inner loop of andvaranaut.
 *
 * The reference (andrew-angus/andvaranaut) has no FFI of its own: its seam is the Python class
 * GPMCMC (andvaranaut/gpmcmc.py:30) and, inside it, three third-party call families that this
 * library replaces one for one:
 *   (i)   the compiled logp(theta) / dlogp(theta) that pm.find_MAP and pm.sample evaluate
 *         (gpmcmc.py:310-323 builds them, :332,:345,:351,:357 consume them)      -> avn_gp_loglik_grad
 *   (ii)  gp.predict(Xnew, point=hypers, diag=True, jitter, pred_noise=True)
 *         (gpmcmc.py:588-598)                                                    -> avn_gp_factorize + avn_gp_predict
 *   (iii) the Gauss-Hermite reversion / expected-improvement loop __gh_stats
 *         (gpmcmc.py:545-569)                                                    -> epilogue of avn_gp_predict
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types; every *_dev pointer is DEVICE memory owned
 *     by the caller; the library allocates nothing on the device and never frees caller memory.
 *     (One exception: a library built with -DAVN_FACTOR_PROF, an instrumentation build that is never shipped,
 *     keeps one small static counter buffer of its own.)
 *   - all work is ordered on the caller's stream (a cudaStream_t passed as void*) and is asynchronous;
 *     only avn_gp_create/avn_gp_destroy/avn_gp_set_data touch no stream.  The library owns a few non-blocking side
 *     streams per handle (sample groups of avn_gp_set_streams, the output-warp column of calls with few samples, the
 *     replayed graph of avn_gp_loglik_grad_host): each is forked from and joined back into the caller's stream by
 *     events inside the call, so the caller sees plain stream order.
 *   - a handle is bound to ONE device: the device that is current when avn_gp_create runs (or, when none is
 *     visible then, when the first call that enqueues work runs).  Every such call makes that device current for
 *     its duration and restores the caller's afterwards, so handles may be used from any host thread and handles
 *     on different devices may share a process; the stream and every *_dev pointer must belong to the handle's
 *     device.  Per-device kernel attributes are set once per handle, not cached process-wide.
 *   - return value: 0 = ok, negative = bad argument / launch failure (text via avn_last_error()).
 *     Numerical failure is DATA, not an error: info[b] = k > 0 means pivot k of sample b was not
 *     positive; then ll[b] = -inf and grad[b,:] = 0 (PyMC's NaN-Cholesky -> -inf convention).
 *     info[b] = -1 means the factorisation was ABORTED: a dataflow wait inside the persistent factor kernel
 *     exceeded its bound (several seconds by default; reachable under time-slicing with another process, a
 *     debugger or a sanitizer).  Then ll[b] = NaN, grad[b,:] = 0 (avn_gp_factorize: the state buffer is invalid)
 *     and the call must be repeated; the bound is set with avn_gp_set_debug.  The calls are asynchronous, so the
 *     abort cannot be a return code: callers that read info must treat a negative value as an error.
 *   - all arithmetic is FP64.
 *
 * Hyperparameter vector theta (constrained space, one row per sample), in this order
 * (names as in GPMCMC.hypers, gpmcmc.py:193-208,217-220,255-264,288):
 *     [gv (only if noise)] [l: d*nkern, kernel-major] [kv: nkern] [iwgp: n_iw] [cw: n_cw, wgp order] [alpha (only if RatQuad)]
 */
#ifndef AVN_GP_H
#define AVN_GP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVN_MAX_D 16        /* input dimensions */
#define AVN_MAX_KERN 4      /* kernels in one sum/product fold */
#define AVN_MAX_STAGES 6    /* stages in one composite warp (wgp) */
#define AVN_MAX_WPARAMS 8   /* learnable parameters of one composite warp */
#define AVN_MAX_GH 32       /* Gauss-Hermite nodes */
#define AVN_TILE 64         /* matrices are padded to a multiple of this */

typedef struct avn_gp avn_gp;

/* pm.gp.cov.* used at gpmcmc.py:283-299 */
enum avn_kernel { AVN_RBF = 0, AVN_MATERN52 = 1, AVN_MATERN32 = 2, AVN_EXPONENTIAL = 3, AVN_RATQUAD = 4 };
/* '+' / '*' of the kernel string, gpmcmc.py:301-307 */
enum avn_kop { AVN_ADD = 0, AVN_MUL = 1 };
/* stages of andvaranaut/transform.py:431-534 */
enum avn_warp_op {
  AVN_W_AFFINE_CONST = 0, AVN_W_AFFINE = 1, AVN_W_LOG = 2, AVN_W_ARCSINH = 3, AVN_W_BOXCOX = 4,
  AVN_W_SINHARCSINH = 5, AVN_W_SAL = 6, AVN_W_KUMARASWAMY = 7, AVN_W_STDSHIFT = 8, AVN_W_MEANSTD = 9,
  AVN_W_MINSHIFT = 10, AVN_W_STDDEV = 11, AVN_W_MAXMIN = 12, AVN_W_PZERO = 13, AVN_W_BOXCOX_CONST = 14
};

typedef struct {
  int32_t op;      /* avn_warp_op */
  int32_t pidx;    /* index of the stage's first learnable parameter inside its warp's parameter slice, or -1: use c[] */
  double c[4];     /* constants (frozen coefficients) */
} avn_warp_stage;

typedef struct {
  int32_t nstages;  /* 0 = identity (column already converted on the host) */
  int32_t nparams;  /* learnable parameters consumed by this warp */
  avn_warp_stage st[AVN_MAX_STAGES];
} avn_warp_prog;

typedef struct {
  int32_t d;                       /* nx */
  int32_t nkern;
  int32_t kern[AVN_MAX_KERN];      /* avn_kernel */
  int32_t op[AVN_MAX_KERN];        /* op[m-1] joins kernel m to the running fold */
  int32_t noise;                   /* 1: theta starts with gv */
  double jitter;
  avn_warp_prog xwarp[AVN_MAX_D];  /* learnable input warps (iwgp=True), consumed in dimension order */
  avn_warp_prog ywarp;             /* learnable output warp (cwgp=True) */
} avn_model_desc;

/* byte offsets of the named buffers inside a loglik workspace (for tests and profiling).
 * fflags: the int32 progress words of the factor kernel -- lflag [B][nb], tflag [B][nb], 64 control words, sflag [B][nb],
 * dflag [B][nb] (the last two only used by launches with few samples) -- followed by the 1026 scheduler words of the
 * single-sample gradient kernel; zeroed by one memset at the start of every call. */
typedef struct {
  int64_t npad, nb;
  int64_t xw, dxw, xs, x2, z, dz, wstat, kl, t, beta, alpha, gpart, gxpart, fpart, fflags, total;
} avn_ws_layout;

/* Gauss-Hermite reversion / expected improvement, gpmcmc.py:545-569 */
typedef struct {
  int32_t mode;        /* 0: return latent mu/var; 1: GH-revert mean/var; 2: EI */
  int32_t deg;
  int32_t normvar;
  int32_t ei_max;      /* EIopt == 'max' */
  double yopt;
  double nodes[AVN_MAX_GH];
  double weights[AVN_MAX_GH];
  avn_warp_prog yrev;  /* frozen output transform; applied right-to-left */
} avn_epilogue;

const char* avn_last_error(void);
int avn_version(void);

int avn_gp_create(const avn_model_desc* desc, avn_gp** out);
void avn_gp_destroy(avn_gp* gp);
int avn_gp_num_params(const avn_gp* gp);

/* training data: X [N,d] row-major (columns with a learnable warp RAW, the others already converted),
 * y [N] (raw, mean-subtracted when the output warp is learnable, else already converted).
 * Registers the two pointers and N in the handle; no device work is enqueued, hence no stream argument (SURVEY 8b
 * sketched one for a copy that this library does not make: it allocates no device memory to copy into).  Lifetime:
 * the caller keeps both arrays alive and unchanged until the next avn_gp_set_data / avn_gp_destroy; every later
 * call reads them on ITS stream, so they must be complete on that stream when it is made. */
int avn_gp_set_data(avn_gp* gp, const double* X_dev, const double* y_dev, int64_t N);

size_t avn_gp_workspace_bytes(const avn_gp* gp, int64_t B);
int avn_gp_workspace_layout(const avn_gp* gp, int64_t B, avn_ws_layout* out);

/* logp / dlogp of the marginal likelihood for B hyperparameter samples at once.
 * theta [B,P]; ll [B]; grad [B,P] (may be NULL: value only); info [B]. */
int avn_gp_loglik_grad(avn_gp* gp, const double* theta_dev, int64_t B, double* ll_dev, double* grad_dev,
                       int32_t* info_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* The same evaluation with HOST buffers -- the call pm.find_MAP / pm.sample make once per optimiser step or leapfrog
 * (gpmcmc.py:332,345,351,357: a NumPy point in, logp and dlogp out).  theta_host [B,P] and out_host [B*(P+2)] are
 * PINNED host memory; out_host receives ll [B], then grad [B,P], then info as int32 [B] in the low half of the last B
 * doubles (one packed device->host copy).  staging_dev: avn_gp_host_staging_bytes(B) bytes of device memory for the
 * points and the packed results.  The host->device copy, every launch and the device->host copy are captured ONCE as
 * a CUDA graph on a library-owned stream (capture is not possible on the legacy default stream) and replayed by later
 * calls with the same buffers, B and data; the replay is ordered behind the work already enqueued on `stream`, and the
 * call returns when out_host is complete (SYNCHRONOUS, unlike the rest of this API).
 * flags: bit 0 = gradient wanted (0: value only); bit 1 = return right after the launch -- the caller does its own host
 * work (the prior terms of the posterior, gpmcmc.py:193-208) meanwhile and completes the call with avn_gp_host_wait;
 * theta_host / out_host must not be touched in between. */
size_t avn_gp_host_staging_bytes(const avn_gp* gp, int64_t B);
int avn_gp_loglik_grad_host(avn_gp* gp, const double* theta_host, int64_t B, double* out_host, int32_t flags,
                            void* staging_dev, size_t staging_bytes, void* ws_dev, size_t ws_bytes, void* stream);
int avn_gp_host_wait(avn_gp* gp);

/* Independent samples of one avn_gp_loglik_grad call are split into up to max_groups (1..8, default 4) groups that
 * run concurrently on library-owned streams, forked from and joined back into the caller's stream. */
int avn_gp_set_streams(avn_gp* gp, int max_groups);

/* covariance build alone (K + (gv+jitter) I, lower block-triangle valid), K [B,npad,npad] */
int avn_gp_cov(avn_gp* gp, const double* theta_dev, int64_t B, double* K_dev, void* ws_dev, size_t ws_bytes,
               void* stream);

/* predict: factorise once for one theta, then stream test points */
size_t avn_gp_state_bytes(const avn_gp* gp);
int avn_gp_factorize(avn_gp* gp, const double* theta_dev, void* state_dev, size_t state_bytes, int32_t* info_dev,
                     void* ws_dev, size_t ws_bytes, void* stream);
size_t avn_gp_predict_workspace_bytes(const avn_gp* gp, int64_t M);
/* Xs [M,d] converted test points; mean_add [M] or NULL (user mean function evaluated on the host);
 * out_mean/out_var [M]. */
int avn_gp_predict(avn_gp* gp, const void* state_dev, const double* Xs_dev, int64_t M, const avn_epilogue* epi,
                   const double* mean_add_dev, double* out_mean_dev, double* out_var_dev, void* ws_dev,
                   size_t ws_bytes, void* stream);

/* BO refine (gpmcmc.py:738-801): the predictive graph the reference differentiates w.r.t. the query point inside
 * pm.find_MAP -- here value AND analytic gradient for M query points at once.  Same epilogue as avn_gp_predict;
 * pred_noise = 0 leaves gv out of the variance as that inline graph does (:775-778), 1 matches gp.predict(pred_noise=True).
 * dmean_add [M,d] = gradient of the user mean function w.r.t. the converted inputs, or NULL.
 * out_dmean / out_dvar [M,d]: d out_mean / d Xs, d out_var / d Xs (converted-input space; the caller chains d con / d x). */
size_t avn_gp_predict_grad_workspace_bytes(const avn_gp* gp, int64_t M);
int avn_gp_predict_grad(avn_gp* gp, const void* state_dev, const double* Xs_dev, int64_t M, const avn_epilogue* epi,
                        int32_t pred_noise, const double* mean_add_dev, const double* dmean_add_dev,
                        double* out_mean_dev, double* out_var_dev, double* out_dmean_dev, double* out_dvar_dev,
                        void* ws_dev, size_t ws_bytes, void* stream);

/* Rank-1 extension of a factorised state by ONE training point with the hyperparameters unchanged: what the data
 * appends of BO / inverse_opt (gpmcmc.py:881-904, :1197-1205) need between two fits when fit_method = 'none' --
 * O(N^2) (four HBM-bound launches, the lower triangle of T read twice) instead of the O(N^3) refactorisation the
 * reference performs inside every predict call (:588-598).
 * xnew [d] converted inputs, znew [1] converted output (both DEVICE memory).  Works in place on state_dev.
 * Returns 1 (nothing done) when the padded slab is full, N % AVN_TILE == 0: the caller then refactorises with
 * N + 1 points.  info[0] = N + 1 if the new pivot is not positive (state unchanged).  The handle is NOT modified:
 * after a successful append the caller registers the N + 1-row arrays with avn_gp_set_data (stream-ordered, the
 * state buffer keeps its size because npad is unchanged). */
size_t avn_gp_append_workspace_bytes(const avn_gp* gp);
int avn_gp_append(avn_gp* gp, void* state_dev, size_t state_bytes, const double* xnew_dev, const double* znew_dev,
                  int32_t* info_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Test / diagnosis knobs of the persistent factor kernel.  wait_bound_log2 in [10,31]: every progress-flag wait gives
 * up after 2^wait_bound_log2 polls (default 26, several seconds) and the call reports info = -1.  fault = 1 injects a
 * fault -- the panel tile P(0,0,1) of sample 0 is never published -- so that tests can exercise that path; 0 = none. */
int avn_gp_set_debug(avn_gp* gp, int wait_bound_log2, int fault);

/* introspection used by bench.py: number of kernel launches issued by the last call on this handle */
int64_t avn_gp_last_launch_count(const avn_gp* gp);

/* phase timing (off by default).  When enabled, avn_gp_loglik_grad / avn_gp_factorize / avn_gp_predict record
 * CUDA events on the caller's stream around each phase; avn_gp_phase_ms synchronises on the last event and
 * returns the elapsed milliseconds of the most recent call, indexed by avn_phase. */
enum avn_phase {
  AVN_PH_WARP = 0, AVN_PH_COV = 1, AVN_PH_FACTOR = 2 /* chol + inverse */, AVN_PH_BETA = 3, AVN_PH_UNUSED4 = 4, AVN_PH_ALPHA = 5,
  AVN_PH_KINV_GRAD = 6, AVN_PH_FINALIZE = 7, AVN_PH_KXS = 8, AVN_PH_PREDICT_VAR = 9, AVN_PH_COUNT = 10
};
int avn_gp_set_profiling(avn_gp* gp, int enable);
int avn_gp_phase_ms(avn_gp* gp, double* out_ms /* [AVN_PH_COUNT] */);

#ifdef __cplusplus
}
#endif
#endif /* AVN_GP_H */

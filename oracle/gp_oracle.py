"""NumPy/SciPy oracle for andvaranaut's GP inner loop (TEST INFRASTRUCTURE ONLY, parity unpinned).

What it restates, and where the reference does it:
  * hyperparameter layout and priors          andvaranaut/gpmcmc.py:193-208,217-220,255-264,288
  * fixed / learnable x and y conversion      gpmcmc.py:211-237, 240-279
  * kernel sum/product fold                   gpmcmc.py:282-307   (PyMC  pm.gp.cov.*, restated below)
  * marginal likelihood, explicit branch      gpmcmc.py:310-319
  * marginal likelihood, Marginal branch      gpmcmc.py:321-323   (PyMC  gp.Marginal + MvNormal.logp)
  * predict(diag=True, pred_noise=True)       gpmcmc.py:588-598   (PyMC  Marginal._build_conditional)
  * Gauss-Hermite reversion / EI              gpmcmc.py:545-569
PyMC 5.9.x formulas (``pymc/gp/cov.py`` Stationary.square_dist / euclidean_dist, ExpQuad, RatQuad,
Matern52, Matern32, Exponential; ``pymc/gp/util.py`` stabilize; ``pymc/distributions/multivariate.py``
quaddist_chol) are restated from the published source -- they cannot be executed here.

Gradients: the reference obtains them by PyTensor reverse-mode autodiff of the expressions above;
the oracle uses the closed forms of the same derivatives (W = alpha alpha^T - K^-1 contractions) and
is itself checked against central finite differences of its own log-likelihood in tests/.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import scipy.linalg as sla

from .warp_oracle import WarpOracle, Dual

KERNELS = ('RBF', 'Matern52', 'Matern32', 'Exponential', 'RatQuad')
SQRT5 = np.sqrt(5.0)
SQRT3 = np.sqrt(3.0)


@dataclass
class ModelSpec:
    """Static description of one GP model (what ``GPMCMC.__fit`` builds inside ``pm.Model()``)."""
    nx: int
    kerns: List[str] = field(default_factory=lambda: ['RBF'])
    ops: List[str] = field(default_factory=list)            # '+' / '*' between consecutive kernels
    noise: bool = True
    jitter: float = 1e-6
    # per input dimension: None (column already converted on the host) or (stage names, (lo, hi) of the prior)
    xwarps: Optional[List[Optional[Tuple[List[str], Optional[Tuple[float, float]]]]]] = None
    ywarp: Optional[List[str]] = None                       # stage names of the learnable output warp

    @property
    def nkern(self):
        return len(self.kerns)

    @property
    def has_alpha(self):
        return 'RatQuad' in self.kerns

    def n_iw(self):
        if self.xwarps is None:
            return 0
        from .warp_oracle import STAGE_PARAMS
        return sum(sum(len(STAGE_PARAMS.get(s, ())) for s in w[0]) for w in self.xwarps if w is not None)

    def n_cw(self):
        if self.ywarp is None:
            return 0
        from .warp_oracle import STAGE_PARAMS
        return sum(len(STAGE_PARAMS.get(s, ())) for s in self.ywarp)

    def offsets(self):
        """Flat constrained-space layout: [gv?][l: nx*nkern][kv: nkern][iwgp][cw (wgp order)][alpha?]."""
        o = {}
        p = 0
        if self.noise:
            o['gv'] = p
            p += 1
        o['l'] = p
        p += self.nx * self.nkern
        o['kv'] = p
        p += self.nkern
        o['iw'] = p
        p += self.n_iw()
        o['cw'] = p
        p += self.n_cw()
        if self.has_alpha:
            o['alpha'] = p
            p += 1
        o['P'] = p
        return o


def unpack(spec: ModelSpec, theta):
    o = spec.offsets()
    theta = np.asarray(theta, dtype=np.float64)
    assert theta.shape == (o['P'],), (theta.shape, o['P'])
    return dict(
        gv=float(theta[o['gv']]) if spec.noise else 0.0,
        l=theta[o['l']:o['l'] + spec.nx * spec.nkern],
        kv=theta[o['kv']:o['kv'] + spec.nkern],
        iw=theta[o['iw']:o['iw'] + spec.n_iw()],
        cw=theta[o['cw']:o['cw'] + spec.n_cw()],
        alpha=float(theta[o['alpha']]) if spec.has_alpha else 1.0,
    )


# ---------------------------------------------------------------------------------------------
# covariance functions (PyMC 5.9 pymc/gp/cov.py)
# ---------------------------------------------------------------------------------------------
def square_dist(X, Xs, ls):
    """Stationary.square_dist: gram form, clipped at 0."""
    X = X * (1.0 / ls)
    X2 = np.sum(np.square(X), 1)
    if Xs is None:
        sqd = -2.0 * np.dot(X, X.T) + (X2.reshape(-1, 1) + X2.reshape(1, -1))
    else:
        Xs = Xs * (1.0 / ls)
        Xs2 = np.sum(np.square(Xs), 1)
        sqd = -2.0 * np.dot(X, Xs.T) + (X2.reshape(-1, 1) + Xs2.reshape(1, -1))
    return np.clip(sqd, 0.0, np.inf)


def kern_from_r2(kind, r2, alpha=1.0):
    """Unit-variance stationary kernel value k and dk/d(r2) given the clipped squared distance.

    ExpQuad: exp(-r2/2).  Matern52/32 and Exponential go through euclidean_dist = sqrt(r2 + 1e-12).
    Exponential is exp(-r/2) (PyMC's convention).  RatQuad: (1 + r2/(2 alpha))^-alpha.
    """
    if kind == 'RBF':
        k = np.exp(-0.5 * r2)
        return k, -0.5 * k
    if kind == 'RatQuad':
        base = 1.0 + 0.5 * r2 * (1.0 / alpha)
        k = np.power(base, -1.0 * alpha)
        return k, -0.5 * np.power(base, -alpha - 1.0)
    r = np.sqrt(r2 + 1e-12)
    if kind == 'Matern52':
        e = np.exp(-1.0 * SQRT5 * r)
        k = (1.0 + SQRT5 * r + 5.0 / 3.0 * np.square(r)) * e
        return k, -(5.0 / 6.0) * (1.0 + SQRT5 * r) * e
    if kind == 'Matern32':
        e = np.exp(-1.0 * SQRT3 * r)
        k = (1.0 + SQRT3 * r) * e
        return k, -1.5 * e
    if kind == 'Exponential':
        k = np.exp(-0.5 * r)
        return k, -k / (4.0 * r)
    raise ValueError(kind)


def ratquad_dalpha(r2, alpha):
    base = 1.0 + 0.5 * r2 / alpha
    k = np.power(base, -alpha)
    return k * (-np.log(base) + (0.5 * r2 / alpha) / base)


def fold_values(spec, vals):
    """Left-to-right fold of the per-kernel matrices (gpmcmc.py:301-307) and the partials
    dK/dvals[k] needed by the product rule."""
    n = len(vals)
    pref = [vals[0]]
    for m in range(1, n):
        pref.append(pref[-1] + vals[m] if spec.ops[m - 1] == '+' else pref[-1] * vals[m])
    K = pref[-1]
    coef = [None] * n
    g = np.ones_like(K)
    for m in range(n - 1, 0, -1):
        if spec.ops[m - 1] == '+':
            coef[m] = g
        else:
            coef[m] = g * pref[m - 1]
            g = g * vals[m]
    coef[0] = g
    return K, coef


def kdiag_total(spec, kv):
    """cov(Xnew, diag=True): every stationary kernel has unit diagonal, folded with the variances."""
    t = kv[0]
    for m in range(1, spec.nkern):
        t = t + kv[m] if spec.ops[m - 1] == '+' else t * kv[m]
    return float(t)


def cov_matrix(spec, th, X, Xs=None, want_parts=False):
    """K (or K_xs) = fold_k kv_k * k_k(X / l_k)."""
    vals, parts = [], []
    for k, kind in enumerate(spec.kerns):
        ls = th['l'][k * spec.nx:(k + 1) * spec.nx]
        r2 = square_dist(X, Xs, ls)
        kk, dk = kern_from_r2(kind, r2, th['alpha'])
        vals.append(th['kv'][k] * kk)
        parts.append((r2, kk, dk))
    K, coef = fold_values(spec, vals)
    if want_parts:
        return K, coef, parts
    return K


# ---------------------------------------------------------------------------------------------
# conversions (host precompute or learnable warps)
# ---------------------------------------------------------------------------------------------
def warp_inputs(spec, th, X, with_duals):
    """xin of gpmcmc.py:224-237.  Columns whose entry in spec.xwarps is None are taken as already
    converted.  Returns Xw [N,nx] and, with duals, dXw [N,nx,P_iw]."""
    N = X.shape[0]
    Piw = spec.n_iw()
    Xw = np.array(X, dtype=np.float64, copy=True)
    dXw = np.zeros((N, spec.nx, Piw)) if with_duals else None
    if spec.xwarps is None:
        return Xw, dXw
    rc = 0
    for i, w in enumerate(spec.xwarps):
        if w is None:
            continue
        names, interval = w
        from .warp_oracle import STAGE_PARAMS
        npar = sum(len(STAGE_PARAMS.get(s, ())) for s in names)
        p = th['iw'][rc:rc + npar]
        wo = WarpOracle(names, p, y=X[:, i], xdist_interval=interval, with_duals=with_duals)
        out = wo._ycon
        if with_duals:
            Xw[:, i] = out.v
            dXw[:, i, rc:rc + npar] = out.d
        else:
            Xw[:, i] = out
        rc += npar
    return Xw, dXw


def warp_outputs(spec, th, y, with_duals):
    """yin, yder of gpmcmc.py:275-279 (y is the raw, mean-subtracted output column)."""
    if spec.ywarp is None:
        return np.asarray(y, dtype=np.float64), None, None, None
    wo = WarpOracle(spec.ywarp, th['cw'], y=y, with_duals=with_duals)
    z = wo._ycon
    if with_duals:
        yd = wo.der(Dual.lift(np.asarray(y, dtype=np.float64), len(th['cw'])))
        return z.v, z.d, yd.v, yd.d
    return z, None, wo.der(np.asarray(y, dtype=np.float64)), None


# ---------------------------------------------------------------------------------------------
# log marginal likelihood and gradient
# ---------------------------------------------------------------------------------------------
@dataclass
class LLResult:
    ll: float
    grad: Optional[np.ndarray]
    info: int
    L: Optional[np.ndarray] = None
    beta: Optional[np.ndarray] = None
    alpha: Optional[np.ndarray] = None
    Xw: Optional[np.ndarray] = None
    z: Optional[np.ndarray] = None


def loglik(spec: ModelSpec, theta, X, y, want_grad=True, keep=False) -> LLResult:
    """log p(y | X, theta) (+ sum log g'(y) when the output warp is learnable) and d/dtheta
    in the flat constrained layout of ``ModelSpec.offsets``."""
    th = unpack(spec, theta)
    o = spec.offsets()
    N = X.shape[0]
    explicit = spec.ywarp is not None
    Xw, dXw = warp_inputs(spec, th, X, with_duals=want_grad and spec.n_iw() > 0)
    z, dz, yder, dyder = warp_outputs(spec, th, y, with_duals=want_grad)
    K, coef, parts = cov_matrix(spec, th, Xw, None, want_parts=True)
    if explicit:
        # gpmcmc.py:312  K += I*(jitter+gvar)
        K = K + np.eye(N) * (spec.jitter + th['gv'])
    else:
        # Marginal: cov + WhiteNoise(sigma=sqrt(gv)) then stabilize(., jitter)
        sig = np.sqrt(th['gv'])
        K = (K + np.eye(N) * np.square(sig)) + spec.jitter * np.eye(N)
    try:
        L = sla.cholesky(K, lower=True, check_finite=False)
    except sla.LinAlgError:
        return LLResult(-np.inf, np.zeros(o['P']) if want_grad else None, 1)
    if not np.all(np.isfinite(np.diag(L))) or np.any(np.diag(L) <= 0):
        return LLResult(-np.inf, np.zeros(o['P']) if want_grad else None, 1)
    beta = sla.solve_triangular(L, z, lower=True, check_finite=False)
    alpha = sla.solve_triangular(L.T, beta, lower=False, check_finite=False)
    if explicit:
        ll = (-0.5 * np.dot(z.T, alpha) - np.sum(np.log(np.diag(L)))
              - 0.5 * N * np.log(2 * np.pi) + np.sum(np.log(yder)))
    else:
        quaddist = np.sum(beta ** 2)
        logdet = np.sum(np.log(np.diag(L)))
        norm = -0.5 * N * np.log(2 * np.pi)
        ll = norm - 0.5 * quaddist - logdet
    res = LLResult(float(ll), None, 0)
    if keep:
        res.L, res.beta, res.alpha, res.Xw, res.z = L, beta, alpha, Xw, z
    if not want_grad:
        return res

    grad = np.zeros(o['P'])
    Kinv = sla.cho_solve((L, True), np.eye(N), check_finite=False)
    W = np.outer(alpha, alpha) - Kinv
    if spec.noise:
        grad[o['gv']] = 0.5 * np.trace(W)
    GX = np.zeros((N, spec.nx)) if spec.n_iw() > 0 else None
    for k, kind in enumerate(spec.kerns):
        r2, kk, dk = parts[k]
        ls = th['l'][k * spec.nx:(k + 1) * spec.nx]
        Wc = W * coef[k]
        grad[o['kv'] + k] = 0.5 * np.sum(Wc * kk)
        # where the gram form was clipped the sub-gradient is zero (pt.clip)
        WK = Wc * th['kv'][k] * dk * (r2 > 0.0)
        for m in range(spec.nx):
            D = Xw[:, m][:, None] - Xw[:, m][None, :]
            # d r2 / d l_m = -2 D^2 / l_m^3 ; dll = 1/2 sum W dK
            grad[o['l'] + k * spec.nx + m] = -np.sum(WK * D * D) / ls[m] ** 3
            if GX is not None:
                GX[:, m] += 2.0 * np.sum(WK * D, axis=1) / ls[m] ** 2
        if kind == 'RatQuad':
            grad[o['alpha']] = 0.5 * np.sum(Wc * th['kv'][k] * ratquad_dalpha(r2, th['alpha']))
    if GX is not None:
        grad[o['iw']:o['iw'] + spec.n_iw()] = np.einsum('nm,nmp->p', GX, dXw)
    if explicit:
        grad[o['cw']:o['cw'] + spec.n_cw()] = -alpha @ dz + np.sum(dyder / yder[:, None], axis=0)
    res.grad = grad
    return res


# ---------------------------------------------------------------------------------------------
# predict + Gauss-Hermite reversion
# ---------------------------------------------------------------------------------------------
def predict(spec: ModelSpec, theta, Xc, z, Xs):
    """Marginal._build_conditional with diag=True, pred_noise=True on already-converted inputs
    (Xc [N,nx], z [N]) and converted test points Xs [M,nx].  Returns (mu [M], var [M])."""
    th = unpack(spec, theta)
    N = Xc.shape[0]
    Kxx = cov_matrix(spec, th, Xc)
    Kxs = cov_matrix(spec, th, Xc, Xs)
    sig2 = np.square(np.sqrt(th['gv']))
    Knx = np.eye(N) * sig2
    L = sla.cholesky((Kxx + spec.jitter * np.eye(N)) + Knx, lower=True, check_finite=False)
    A = sla.solve_triangular(L, Kxs, lower=True, check_finite=False)
    v = sla.solve_triangular(L, z, lower=True, check_finite=False)
    mu = np.dot(A.T, v)
    Kss = kdiag_total(spec, th['kv']) * np.ones(Xs.shape[0])
    var = Kss - np.sum(np.square(A), 0)
    var = var + sig2
    return mu, var


def predict_blocked(spec, theta, Xc, z, Xs, block=4096, L=None):
    """Same arithmetic as :func:`predict`, with L factorised once and test points streamed in
    blocks (used by the CPU baseline; the reference itself refactorises on every call)."""
    th = unpack(spec, theta)
    N = Xc.shape[0]
    sig2 = np.square(np.sqrt(th['gv']))
    if L is None:
        Kxx = cov_matrix(spec, th, Xc)
        L = sla.cholesky((Kxx + spec.jitter * np.eye(N)) + np.eye(N) * sig2, lower=True, check_finite=False)
    v = sla.solve_triangular(L, z, lower=True, check_finite=False)
    kd = kdiag_total(spec, th['kv'])
    mu = np.empty(Xs.shape[0])
    var = np.empty(Xs.shape[0])
    for s in range(0, Xs.shape[0], block):
        Kxs = cov_matrix(spec, th, Xc, Xs[s:s + block])
        A = sla.solve_triangular(L, Kxs, lower=True, check_finite=False)
        mu[s:s + block] = np.dot(A.T, v)
        var[s:s + block] = kd - np.sum(np.square(A), 0) + sig2
    return mu, var, L


def predict_grad(spec: ModelSpec, theta, Xc, z, Xs, pred_noise=True):
    """Latent mean / variance and their gradients w.r.t. the converted query points Xs [M,nx]: the expression the
    reference builds inline for the BO refine step and hands to PyTensor autodiff (gpmcmc.py:738-778: kstar,
    v = solve_triangular(L, kstar), ycpmean = kstar^T alpha, ycpvar = k** - v^T v; note: no ``+ gv`` there, i.e.
    ``pred_noise=False``).  Closed forms of the same derivatives:
        d mu / d x = sum_i alpha_i dk_i/dx,   d var / d x = -2 sum_i (K^-1 k*)_i dk_i/dx,
        dk_i/dx_m = sum_q coef_q kv_q k_q'(r2_q) * 2 (x_m - X_im) / l_qm^2   (zero where square_dist was clipped).
    Returns (mu [M], var [M], dmu [M,nx], dvar [M,nx])."""
    th = unpack(spec, theta)
    N, M = Xc.shape[0], Xs.shape[0]
    Kxx = cov_matrix(spec, th, Xc)
    sig2 = np.square(np.sqrt(th['gv']))
    L = sla.cholesky((Kxx + spec.jitter * np.eye(N)) + np.eye(N) * sig2, lower=True, check_finite=False)
    Kxs, coef, parts = cov_matrix(spec, th, Xc, Xs, want_parts=True)
    A = sla.solve_triangular(L, Kxs, lower=True, check_finite=False)
    v = sla.solve_triangular(L, z, lower=True, check_finite=False)
    alpha = sla.solve_triangular(L.T, v, lower=False, check_finite=False)
    Wm = sla.solve_triangular(L.T, A, lower=False, check_finite=False)      # K^-1 K_xs  [N,M]
    mu = np.dot(A.T, v)
    var = kdiag_total(spec, th['kv']) - np.sum(np.square(A), 0)
    if pred_noise:
        var = var + sig2
    dmu = np.zeros((M, spec.nx))
    dvar = np.zeros((M, spec.nx))
    for k in range(spec.nkern):
        r2, kk, dk = parts[k]
        ls = th['l'][k * spec.nx:(k + 1) * spec.nx]
        F = coef[k] * th['kv'][k] * dk * (r2 > 0.0)                          # [N,M]
        for m in range(spec.nx):
            D = (Xs[:, m][None, :] - Xc[:, m][:, None]) * (2.0 / ls[m] ** 2)  # d r2 / d x_m
            dmu[:, m] += np.sum(alpha[:, None] * F * D, axis=0)
            dvar[:, m] += -2.0 * np.sum(Wm * F * D, axis=0)
    return mu, var, dmu, dvar


def inverse_loglik(spec: ModelSpec, theta, Xc, yc, xo_c, yo_c, ynoise, log_yder_sum=0.0):
    """Potential of the Bayesian inverse problem, literally as the reference builds it (gpmcmc.py:1098-1165):
    the unknown converted point ``xo_c`` [nx] is appended ``nobs = len(yo_c)`` times to the converted training inputs
    (:1099-1104, the scalar x prior broadcast into every observation row), the kernel is evaluated on the stacked inputs
    with the FITTED hyperparameters (:1106-1130), ``ynoise`` [N+nobs] is added to the DIAGONAL as is -- the reference
    stores square roots of variances in it, quirk 8 of SURVEY app. C (:1137-1158) -- and
    ``-1/2 y^T K^-1 y - sum log diag L - n/2 log 2 pi + sum log yder`` follows (:1156-1165).  ``spec.jitter`` is not
    used here: the jitter sits inside ``ynoise``."""
    th = unpack(spec, theta)
    nobs = len(yo_c)
    xin = np.vstack([Xc, np.tile(np.asarray(xo_c, dtype=np.float64)[None, :], (nobs, 1))])
    yin = np.r_[yc, yo_c]
    K = cov_matrix(spec, th, xin)
    K = K + np.diag(ynoise)
    L = sla.cholesky(K, lower=True, check_finite=False)
    beta = sla.solve_triangular(L, yin, lower=True, check_finite=False)
    alpha = sla.solve_triangular(L.T, beta, lower=False, check_finite=False)
    return (-0.5 * np.dot(yin.T, alpha) - np.sum(np.log(np.diag(L))) - 0.5 * len(yin) * np.log(2 * np.pi)
            + log_yder_sum)


def gh_stats_inv(y, yv, con, deg=8):
    """``__gh_stats_inv`` (gpmcmc.py:573-585): Gauss-Hermite variance of the CONVERTED observation; the loop
    overwrites ``yvcon`` so the value of the LAST observation is what comes back (a scalar)."""
    xi, wi = np.polynomial.hermite.hermgauss(deg)
    yvcon = None
    for i in range(len(y)):
        yi = np.sqrt(2 * yv[i, 0]) * xi + y[i, 0]
        yir = con(yi)
        ym = 1 / np.sqrt(np.pi) * np.sum(wi * yir)
        yir2 = np.power(yir, 2)
        ym2 = 1 / np.sqrt(np.pi) * np.sum(wi * yir2)
        yvcon = ym2 - ym ** 2
    return yvcon


def gh_stats_grad(mu, var, dmu, dvar, rev, der, mean_add=None, dmean_add=None, normvar=True, deg=8, EI=False,
                  EIopt=None, yopt=None):
    """:func:`gh_stats` together with the gradient of both outputs w.r.t. the query points, given the gradients of
    the latent mean / variance (chain rule of gpmcmc.py:780-801).  ``rev`` is the frozen output reversion, ``der``
    the derivative of its forward map (``d con / d y``), so ``d rev / d z = 1 / der(rev(z))``."""
    xi, wi = np.polynomial.hermite.hermgauss(deg)
    mu = np.asarray(mu, dtype=np.float64).reshape(-1)
    var = np.asarray(var, dtype=np.float64).reshape(-1)
    sd = np.sqrt(2 * var)
    yi = sd[:, None] * xi[None, :] + mu[:, None]
    y0 = rev(yi)
    dyr = 1.0 / der(y0)
    yir = y0 if mean_add is None else y0 + np.asarray(mean_add).reshape(-1, 1)
    dyv = dyr * xi[None, :] / sd[:, None]
    if EI:
        dff = yir - yopt if EIopt == 'max' else yopt - yir
        f = np.where(dff > 0.0, dff, 0.0)
        df = np.where(dff > 0.0, 1.0 if EIopt == 'max' else -1.0, 0.0)
    else:
        f, df = yir, np.ones_like(yir)
    c = 1 / np.sqrt(np.pi)
    m = c * np.sum(wi * f, axis=1)
    m2 = c * np.sum(wi * yir ** 2, axis=1)
    v = m2 - m ** 2
    am, av, cm = c * np.sum(wi * df * dyr, axis=1), c * np.sum(wi * df * dyv, axis=1), c * np.sum(wi * df, axis=1)
    bm = c * np.sum(wi * 2 * yir * dyr, axis=1) - 2 * m * am
    bv = c * np.sum(wi * 2 * yir * dyv, axis=1) - 2 * m * av
    cv = c * np.sum(wi * 2 * yir, axis=1) - 2 * m * cm
    if normvar:
        bm, bv, cv = (b / m ** 2 - 2 * v * a / m ** 3 for a, b in ((am, bm), (av, bv), (cm, cv)))
        v = v / m ** 2
    dma = np.zeros_like(dmu) if dmean_add is None else np.asarray(dmean_add)
    dm_out = am[:, None] * dmu + av[:, None] * dvar + cm[:, None] * dma
    dv_out = bm[:, None] * dmu + bv[:, None] * dvar + cv[:, None] * dma
    return m, v, dm_out, dv_out


def gh_stats_loop(mu, var, rev, mean_add=None, normvar=True, deg=8, EI=False, EIopt=None, yopt=None):
    """Literal restatement of GPMCMC.__gh_stats (gpmcmc.py:545-569): per-point Python loop."""
    xi, wi = np.polynomial.hermite.hermgauss(deg)
    y = np.array(mu, dtype=np.float64).reshape(-1, 1).copy()
    yv = np.array(var, dtype=np.float64).reshape(-1, 1).copy()
    for i in range(len(y)):
        yi = np.sqrt(2 * yv[i, 0]) * xi + y[i, 0]
        yir = rev(yi) + (0.0 if mean_add is None else mean_add[i])
        if EI:
            ydiff = yir - yopt if EIopt == 'max' else yopt - yir
            ydiff = np.where(ydiff > 0.0, ydiff, 0.0)
            y[i, 0] = 1 / np.sqrt(np.pi) * np.sum(wi * ydiff)
        else:
            y[i, 0] = 1 / np.sqrt(np.pi) * np.sum(wi * yir)
        yir2 = np.power(yir, 2)
        ym2 = 1 / np.sqrt(np.pi) * np.sum(wi * yir2)
        yv[i, 0] = ym2 - y[i, 0] ** 2
    if normvar:
        yv /= np.power(y, 2)
    return y, yv


def gh_stats(mu, var, rev, mean_add=None, normvar=True, deg=8, EI=False, EIopt=None, yopt=None):
    """Vectorised form of :func:`gh_stats_loop` (same sums, node-major)."""
    xi, wi = np.polynomial.hermite.hermgauss(deg)
    mu = np.asarray(mu, dtype=np.float64).reshape(-1)
    var = np.asarray(var, dtype=np.float64).reshape(-1)
    yi = np.sqrt(2 * var)[:, None] * xi[None, :] + mu[:, None]
    yir = rev(yi)
    if mean_add is not None:
        yir = yir + np.asarray(mean_add).reshape(-1, 1)
    if EI:
        ydiff = yir - yopt if EIopt == 'max' else yopt - yir
        ydiff = np.where(ydiff > 0.0, ydiff, 0.0)
        m = 1 / np.sqrt(np.pi) * np.sum(wi[None, :] * ydiff, axis=1)
    else:
        m = 1 / np.sqrt(np.pi) * np.sum(wi[None, :] * yir, axis=1)
    m2 = 1 / np.sqrt(np.pi) * np.sum(wi[None, :] * np.power(yir, 2), axis=1)
    v = m2 - m ** 2
    if normvar:
        v = v / np.power(m, 2)
    return m.reshape(-1, 1), v.reshape(-1, 1)


# ---------------------------------------------------------------------------------------------
# priors (PyMC logp formulas; gpmcmc.py:193-208,217-220,255-264,288), checked in tests against scipy.stats
# ---------------------------------------------------------------------------------------------
def logp_lognormal(x, mu, sigma):
    return -0.5 * ((np.log(x) - mu) / sigma) ** 2 - np.log(sigma) - 0.5 * np.log(2 * np.pi) - np.log(x)


def logp_halfnormal(x, sigma):
    return -0.5 * (x / sigma) ** 2 + 0.5 * np.log(2 / np.pi) - np.log(sigma)


def logp_normal(x, mu, sigma):
    return -0.5 * ((x - mu) / sigma) ** 2 - np.log(sigma) - 0.5 * np.log(2 * np.pi)


def logp_truncnormal(x, mu, sigma, lower, upper):
    from scipy.stats import norm
    a, b = (lower - mu) / sigma, (upper - mu) / sigma
    return logp_normal(x, mu, sigma) - np.log(norm.cdf(b) - norm.cdf(a))

"""Extended-precision truth for the GP arithmetic (TEST INFRASTRUCTURE ONLY).

``oracle/gp_oracle.py`` restates PyMC's float64 formulas; this module evaluates the SAME formulas (gpmcmc.py:282-323
and :588-598 through pm.gp.cov.* / gp.Marginal, as listed in gp_oracle's header) in 50-digit arithmetic (mpmath) on
the float64 inputs taken as exact numbers.  In exact arithmetic the gram form of ``Stationary.square_dist`` equals the
direct form and its clip never acts, so the truth uses the direct form; the ``+ 1e-12`` under the square root of
``euclidean_dist`` is part of the function and is kept.  The result is what both the NumPy oracle and the CUDA path
approximate, so it MEASURES the rounding error of each: on the ill-conditioned tutorial case (C1: RBF, noise=False,
jitter 1e-6, cond(K) ~ 6e9) the parity tests compare the device error with the oracle's own error instead of asserting
a blanket ``cond * eps``.

Scope: single stationary kernel (any of the five), fixed conversions (no learnable warps), optional noise -- the
models of C1, C3 and C4.  Pure-Python loops over mpmath numbers: N = 100 takes a few seconds per hyperparameter
vector; meant for generating the committed fixtures (tests/golden/make_truth.py), not for use at test time.
"""
import numpy as np

try:
    import mpmath as mp
except ImportError:  # pragma: no cover - mpmath ships with the image (a torch dependency)
    mp = None

DPS = 50


def _mpf_rows(a):
    return [[mp.mpf(float(v)) for v in row] for row in np.atleast_2d(np.asarray(a, dtype=np.float64))]


def _kern(kind, r2, alpha):
    """unit-variance kernel value k and dk/d(r2) (pymc/gp/cov.py formulas, see gp_oracle.kern_from_r2)."""
    if kind == 'RBF':
        k = mp.exp(-r2 / 2)
        return k, -k / 2
    if kind == 'RatQuad':
        base = 1 + r2 / (2 * alpha)
        return base ** (-alpha), -(base ** (-alpha - 1)) / 2
    r = mp.sqrt(r2 + mp.mpf(1e-12))
    s5, s3 = mp.sqrt(5), mp.sqrt(3)
    if kind == 'Matern52':
        e = mp.exp(-s5 * r)
        return (1 + s5 * r + mp.mpf(5) / 3 * r * r) * e, -(mp.mpf(5) / 6) * (1 + s5 * r) * e
    if kind == 'Matern32':
        e = mp.exp(-s3 * r)
        return (1 + s3 * r) * e, -mp.mpf(3) / 2 * e
    if kind == 'Exponential':
        k = mp.exp(-r / 2)
        return k, -k / (4 * r)
    raise ValueError(kind)


def _cholesky(K):
    n = len(K)
    L = [[mp.mpf(0)] * n for _ in range(n)]
    for j in range(n):
        s = K[j][j] - mp.fsum(L[j][k] * L[j][k] for k in range(j))
        if s <= 0:
            raise ValueError(f'not positive definite at pivot {j + 1}')
        L[j][j] = mp.sqrt(s)
        inv = 1 / L[j][j]
        for i in range(j + 1, n):
            L[i][j] = (K[i][j] - mp.fsum(L[i][k] * L[j][k] for k in range(j))) * inv
    return L


def _solve_lower(L, b):
    n = len(L)
    x = [mp.mpf(0)] * n
    for i in range(n):
        x[i] = (b[i] - mp.fsum(L[i][k] * x[k] for k in range(i))) / L[i][i]
    return x


def _solve_upper_t(L, b):
    """L^T x = b"""
    n = len(L)
    x = [mp.mpf(0)] * n
    for i in range(n - 1, -1, -1):
        x[i] = (b[i] - mp.fsum(L[k][i] * x[k] for k in range(i + 1, n))) / L[i][i]
    return x


def evaluate(kind, noise, jitter, theta, X, z, Xs=None, want_grad=True):
    """theta = [gv (if noise)] [l: d] [kv] [alpha (RatQuad)] on converted inputs X [N,d], converted outputs z [N].
    Returns a dict of float64 arrays rounded from the 50-digit values: ll, grad (layout of theta), and, with Xs [M,d],
    mu / var of predict(diag=True, pred_noise=True); plus cond_est = (max L_ii / min L_ii)^2."""
    if mp is None:
        raise ImportError('mpmath is needed to generate the extended-precision fixtures')
    with mp.workdps(DPS):
        theta = [mp.mpf(float(t)) for t in np.asarray(theta, dtype=np.float64)]
        Xr = _mpf_rows(X)
        n, d = len(Xr), len(Xr[0])
        p = 0
        gv = mp.mpf(0)
        if noise:
            gv = theta[0]
            p = 1
        ls = theta[p:p + d]
        kv = theta[p + d]
        alpha = theta[p + d + 1] if kind == 'RatQuad' else mp.mpf(1)
        zz = [mp.mpf(float(v)) for v in np.asarray(z, dtype=np.float64)]
        jit = mp.mpf(float(jitter))
        r2 = [[mp.fsum(((Xr[i][m] - Xr[j][m]) / ls[m]) ** 2 for m in range(d)) for j in range(n)] for i in range(n)]
        kk = [[None] * n for _ in range(n)]
        dk = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(i + 1):
                kk[i][j], dk[i][j] = _kern(kind, r2[i][j], alpha)
                kk[j][i], dk[j][i] = kk[i][j], dk[i][j]
        K = [[kv * kk[i][j] + ((gv + jit) if i == j else 0) for j in range(n)] for i in range(n)]
        L = _cholesky(K)
        beta = _solve_lower(L, zz)
        logdet = mp.fsum(mp.log(L[i][i]) for i in range(n))
        ll = -mp.mpf(n) / 2 * mp.log(2 * mp.pi) - mp.fsum(b * b for b in beta) / 2 - logdet
        diag = [L[i][i] for i in range(n)]
        out = dict(ll=np.float64(ll), cond_est=np.float64((max(diag) / min(diag)) ** 2))
        alpha_v = _solve_upper_t(L, beta)
        if want_grad:
            # K^-1 = T^T T with T = L^-1 (column by column)
            T = [[mp.mpf(0)] * n for _ in range(n)]
            for c in range(n):
                e = [mp.mpf(1) if i == c else mp.mpf(0) for i in range(n)]
                col = _solve_lower(L, e)
                for i in range(n):
                    T[i][c] = col[i]
            W = [[alpha_v[i] * alpha_v[j] - mp.fsum(T[k][i] * T[k][j] for k in range(max(i, j), n)) for j in range(n)]
                 for i in range(n)]
            grad = []
            if noise:
                grad.append(mp.fsum(W[i][i] for i in range(n)) / 2)
            for m in range(d):
                s = mp.fsum(W[i][j] * kv * dk[i][j] * (Xr[i][m] - Xr[j][m]) ** 2 for i in range(n) for j in range(n))
                grad.append(-s / ls[m] ** 3)
            grad.append(mp.fsum(W[i][j] * kk[i][j] for i in range(n) for j in range(n)) / 2)
            if kind == 'RatQuad':
                def dal(r):
                    base = 1 + r / (2 * alpha)
                    return base ** (-alpha) * (-mp.log(base) + (r / (2 * alpha)) / base)
                grad.append(mp.fsum(W[i][j] * kv * dal(r2[i][j]) for i in range(n) for j in range(n)) / 2)
            out['grad'] = np.array([np.float64(g) for g in grad])
        if Xs is not None and len(Xs):
            Xq = _mpf_rows(Xs)
            mu, var = [], []
            for q in Xq:
                ks = [kv * _kern(kind, mp.fsum(((Xr[i][m] - q[m]) / ls[m]) ** 2 for m in range(d)), alpha)[0]
                      for i in range(n)]
                a = _solve_lower(L, ks)
                mu.append(mp.fsum(ai * bi for ai, bi in zip(a, beta)))
                var.append(kv - mp.fsum(ai * ai for ai in a) + gv)
            out['mu'] = np.array([np.float64(v) for v in mu])
            out['var'] = np.array([np.float64(v) for v in var])
        return out

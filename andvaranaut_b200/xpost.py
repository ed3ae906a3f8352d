"""Posteriors over an unknown INPUT point x with the hyperparameters held at their fitted values: the two PyMC
models the reference builds over ``x0 .. x{nx-1}`` and hands to ``pm.find_MAP`` / ``pm.sample``:

  * the Bayesian inverse problem ``GPMCMC.inverse_opt`` (andvaranaut/gpmcmc.py:1040-1217): given observations
    ``yobs``, the unknown point is appended to the training set and the marginal likelihood of the enlarged data set
    is the potential;
  * the BO ``opt_method='map'`` acquisition model (gpmcmc.py:699-858): potential = acquisition value.

Both expose the interface of :class:`andvaranaut_b200.drivers.Posterior` (``space``, ``logp_dlogp``) so the same
MAP / lock-step MCMC drivers run them, every call evaluating a whole batch of candidate points (restarts, chains)
with ONE ``avn_gp_predict_grad`` launch sequence on the device.

Inverse problem on the device.  The reference refactorises the (N+nobs)x(N+nobs) matrix for every x.  With the
hyperparameters fixed only the last nobs rows depend on x, so the block form is used instead (identical value):

    K = [[A, k 1^T], [1 k^T, c 11^T + D_o]],  A = K_tt + diag(ynoise_t)  (factorised ONCE, ``avn_gp_factorize``)
    mu = k^T A^-1 y_t,  s = c - k^T A^-1 k    (``avn_gp_predict_grad``, latent epilogue, no noise term)
    S  = s 11^T + D_o,  e = y_o - mu 1
    logp(x) = C_train - 1/2 e^T S^-1 e - 1/2 log det S - nobs/2 log 2 pi + sum log yder

with gradient  d logp = (1^T S^-1 e) d mu + 1/2 ((1^T S^-1 e)^2 - 1^T S^-1 1) d s  chained through the device's
d mu / d x, d s / d x and the input conversions.  ``c`` is the diagonal of the FULL-form kernel matrix, i.e. the
kernels evaluated at r = sqrt(0 + 1e-12), not the ``diag=True`` value ``avn_gp_predict`` uses:
the host adds the (hyperparameter-only) difference.
"""
import numpy as np

from .priors import XSpace

SQRT5, SQRT3 = np.sqrt(5.0), np.sqrt(3.0)
HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


def kdiag_values(kerns, ops, kv, alpha=1.0):
    """(diagonal of the full-form kernel matrix, ``diag=True`` value): the left-to-right fold (gpmcmc.py:301-307) of
    kv_k k_k(r2 = 0) with PyMC's euclidean distance sqrt(r2 + 1e-12), and of kv_k alone."""
    r = np.sqrt(1e-12)
    unit = {'RBF': 1.0, 'RatQuad': 1.0,
            'Matern52': (1.0 + SQRT5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-1.0 * SQRT5 * r),
            'Matern32': (1.0 + SQRT3 * r) * np.exp(-1.0 * SQRT3 * r),
            'Exponential': np.exp(-0.5 * r)}
    kv = np.asarray(kv, dtype=np.float64).reshape(-1)
    full, diag = kv[0] * unit[kerns[0]], kv[0]
    for m in range(1, len(kerns)):
        v = kv[m] * unit[kerns[m]]
        full, diag = (full + v, diag + kv[m]) if ops[m - 1] == '+' else (full * v, diag * kv[m])
    return float(full), float(diag)


class XPosterior:
    """prior over x (scipy priors -> PyMC variables, :class:`XSpace`) + a potential evaluated for a batch of raw
    points: ``potential(x [B,nx]) -> (value [B], d value / d x [B,nx])``."""

    def __init__(self, priors, potential):
        self.space = XSpace(priors)
        self.potential = potential
        self.n_eval = self.n_calls = 0

    def logp_dlogp(self, z, jacobian):
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        x, dxdz, ljac, dljac = self.space.theta_from_z(z)
        with np.errstate(all='ignore'):
            val, gx = self.potential(x)
            lp, glp = self.space.prior(x)
            val = val + lp
            grad = self.space.grad_theta_to_z(gx + glp, dxdz)
        self.n_eval += z.shape[0]
        self.n_calls += 1
        if jacobian:
            val = val + ljac
            grad = grad + dljac
        bad = ~np.isfinite(val) | ~np.all(np.isfinite(grad), axis=1)
        val = np.where(bad, -np.inf, val)
        grad[bad] = 0.0
        return val, grad, bad.astype(np.int32)


class InverseLikelihood:
    """potential of the inverse problem (see the module docstring).  ``engine`` is a GPEngine over (xc, yc) factorised
    with the training part of ``ynoise`` on the diagonal; ``con_with_der`` maps raw x [B,nx] to (converted x, d con / d x)."""

    def __init__(self, engine, con_with_der, yo_c, noise_o, c_shift, const):
        self.engine, self.con_with_der = engine, con_with_der
        self.yo = np.asarray(yo_c, dtype=np.float64).reshape(-1)
        self.noise_o = np.broadcast_to(np.asarray(noise_o, dtype=np.float64), self.yo.shape).copy()
        self.c_shift, self.const = float(c_shift), float(const)

    def __call__(self, x):
        xc, dc = self.con_with_der(x)
        mu, s, dmu, ds = (t.cpu().numpy() for t in self.engine.predict_grad(xc, pred_noise=False))
        s = s + self.c_shift
        nobs = len(self.yo)
        B = len(x)
        e = self.yo[None, :] - mu[:, None]
        val = np.full(B, -np.inf)
        gmu = np.zeros(B)
        gs = np.zeros(B)
        # S = s 11^T + D_o per candidate (nobs x nobs, nobs is 1 in the reference's own use): batched Cholesky of the
        # candidates whose S is finite with positive pivots; the others keep logp = -inf
        S = s[:, None, None] * np.ones((nobs, nobs)) + np.diag(self.noise_o)[None, :, :]
        ok = np.isfinite(S).all(axis=(1, 2)) & np.isfinite(e).all(axis=1)
        if nobs == 1:
            ok &= np.where(ok, S[:, 0, 0], 0.0) > 0.0
            d0 = S[ok, 0, 0]
            Se, S1 = e[ok, 0] / d0, 1.0 / d0
            val[ok] = self.const - 0.5 * e[ok, 0] * Se - 0.5 * np.log(d0) - HALF_LOG_2PI
            gmu[ok] = Se
            gs[ok] = 0.5 * (Se ** 2 - S1)
        else:
            for b in np.where(ok)[0]:
                try:
                    Lb = np.linalg.cholesky(S[b])
                except np.linalg.LinAlgError:
                    continue
                u = np.linalg.solve(Lb, np.stack([e[b], np.ones(nobs)], axis=1))
                Se, S1 = np.linalg.solve(Lb.T, u).T
                val[b] = self.const - 0.5 * np.dot(e[b], Se) - np.sum(np.log(np.diag(Lb))) - nobs * HALF_LOG_2PI
                gmu[b] = np.sum(Se)
                gs[b] = 0.5 * (np.sum(Se) ** 2 - np.sum(S1))
        gx = (gmu[:, None] * dmu + gs[:, None] * ds) * dc
        return val, gx

"""MAP and MCMC drivers: the callers that turn batched logp/dlogp evaluations into ``fit`` results.

They stand where ``pm.find_MAP`` (andvaranaut/gpmcmc.py:332,345,357) and ``pm.sample`` (:351) stand in the
reference, and consume the same quantity -- log posterior and its gradient over the unconstrained variables --
but evaluated for MANY hyperparameter vectors per device call:
  * ``find_map``        SciPy L-BFGS-B exactly as PyMC drives it (objective -logp without transform Jacobian,
                        ``maxeval``, non-finite values mapped to a large penalty);
  * ``find_map_multi``  several independent L-BFGS-B runs (restarts) advanced in lock-step so every round of
                        objective evaluations is ONE batched call;
  * ``sample``          B chains of Hamiltonian Monte Carlo advanced in lock-step (one batched call per leapfrog
                        step), dual-averaging step size and diagonal mass adaptation during tuning.  The
                        reference uses PyMC's NUTS with one process per chain; a lock-step NUTS is a "next" row
                         -- the sampled density (logp WITH Jacobian) is identical.
"""
import threading

import numpy as np
from scipy.optimize import minimize

BIG = 1.0e100


class Posterior:
    """log posterior over the unconstrained vector z for one model = GP marginal likelihood on the device
    + priors / transforms on the host."""

    def __init__(self, engine, space, shard=None):
        self.engine, self.space, self.shard = engine, space, shard
        self.n_eval = 0
        self.n_calls = 0

    def logp_dlogp(self, z, jacobian):
        """z [B,P] -> (logp [B], dlogp/dz [B,P], info [B]).  Samples whose covariance is not positive definite
        get logp = -inf and a zero gradient."""
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        theta, dxdz, ljac, dljac = self.space.theta_from_z(z)
        if self.shard is not None:
            ll, gll, info = self.shard.loglik_grad(self.engine, theta)
        else:
            ll_t, g_t, info_t = self.engine.loglik_grad(theta)
            ll, gll, info = ll_t.cpu().numpy(), g_t.cpu().numpy(), info_t.cpu().numpy()
        self.n_eval += z.shape[0]
        self.n_calls += 1
        lp, glp = self.space.prior(theta)
        val = ll + lp
        with np.errstate(all='ignore'):
            grad = self.space.grad_theta_to_z(gll + glp, dxdz)
        if jacobian:
            val = val + ljac
            grad = grad + dljac
        bad = (info != 0) | ~np.isfinite(val)
        val = np.where(bad, -np.inf, val)
        grad[bad] = 0.0
        return val, grad, info


def find_map(post, z0, maxeval=5000, method='L-BFGS-B', **kwargs):
    """One L-BFGS-B run from z0, mirroring pm.find_MAP (pymc/tuning/starting.py): minimise -logp(jacobian=False)."""
    count = [0]

    def cost(z):
        count[0] += 1
        if count[0] > maxeval:
            raise StopIteration
        v, g, _ = post.logp_dlogp(z[None, :], jacobian=False)
        if not np.isfinite(v[0]):
            return BIG, np.zeros_like(z)
        return -v[0], -g[0]

    best = {'z': np.array(z0, dtype=np.float64)}
    try:
        res = minimize(cost, np.array(z0, dtype=np.float64), method=method, jac=True, **kwargs)
        best['z'] = res.x
    except StopIteration:
        pass
    v, _, _ = post.logp_dlogp(best['z'][None, :], jacobian=False)
    return best['z'], float(v[0]), count[0]


class _Rendezvous:
    """Lets R optimiser threads submit one point each and serves them with a single batched evaluation."""

    def __init__(self, post, n):
        self.post, self.active = post, n
        self.cv = threading.Condition()
        self.pending = {}
        self.results = {}
        self.gen = 0

    def _flush(self):
        ids = sorted(self.pending)
        z = np.stack([self.pending[i] for i in ids])
        v, g, _ = self.post.logp_dlogp(z, jacobian=False)
        for k, i in enumerate(ids):
            self.results[i] = (v[k], g[k])
        self.pending.clear()
        self.gen += 1
        self.cv.notify_all()

    def evaluate(self, i, z):
        with self.cv:
            self.pending[i] = np.array(z, dtype=np.float64)
            if len(self.pending) == self.active:
                self._flush()
            else:
                gen = self.gen
                while self.gen == gen:
                    self.cv.wait()
            return self.results.pop(i)

    def leave(self):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) == self.active:
                self._flush()


def find_map_multi(post, z0s, maxeval=5000, method='L-BFGS-B', **kwargs):
    """R independent L-BFGS-B runs; every round of objective evaluations is one batched device call.
    Returns (z [R,P], logp [R])."""
    z0s = np.atleast_2d(np.asarray(z0s, dtype=np.float64))
    R = z0s.shape[0]
    if R == 1:
        z, v, _ = find_map(post, z0s[0], maxeval=maxeval, method=method, **kwargs)
        return z[None, :], np.array([v])
    rv = _Rendezvous(post, R)
    out = [None] * R

    def run(i):
        n = [0]

        def cost(z):
            n[0] += 1
            if n[0] > maxeval:
                raise StopIteration
            v, g = rv.evaluate(i, z)
            if not np.isfinite(v):
                return BIG, np.zeros_like(z)
            return -v, -g
        zi = z0s[i].copy()
        try:
            zi = minimize(cost, zi, method=method, jac=True, **kwargs).x
        except StopIteration:
            pass
        except Exception as e:  # a failed restart must not block the others (gpmcmc.py:337-339)
            print('Restart failed', e)
        finally:
            rv.leave()
        out[i] = zi

    threads = [threading.Thread(target=run, args=(i,)) for i in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    zs = np.stack(out)
    v, _, _ = post.logp_dlogp(zs, jacobian=False)
    return zs, v


class Trace:
    """Minimal stand-in for the ArviZ InferenceData the reference reads (gpmcmc.py:404-430):
    ``posterior[name]`` arrays of shape [chain, draw, ...] and ``sample_stats['lp']`` [chain, draw]."""

    def __init__(self, posterior, lp, z, accept, step_size):
        self.posterior = posterior
        self.sample_stats = {'lp': lp, 'acceptance_rate': accept, 'step_size': step_size}
        self.z = z


def sample(post, draws=1000, tune=1000, chains=4, seed=None, target_accept=0.8, path_length=None,
           max_leapfrog=64, start_z=None, init_jitter=1.0, progressbar=False):
    """Lock-step HMC over ``chains`` chains.  Every leapfrog step is one batched logp/dlogp call."""
    rng = np.random.default_rng(seed)
    sp = post.space
    P = sp.P
    z = (np.tile(sp.initial_z(), (chains, 1)) if start_z is None else np.atleast_2d(start_z).copy())
    if z.shape[0] != chains:
        z = np.tile(z[0], (chains, 1))
    z = z + init_jitter * rng.uniform(-1, 1, size=z.shape)      # PyMC "jitter+adapt_diag" initialisation
    lp, g, _ = post.logp_dlogp(z, True)
    for _ in range(20):                                          # re-draw chains that start at a non-PD point
        bad = ~np.isfinite(lp)
        if not bad.any():
            break
        z[bad] = np.tile(sp.initial_z(), (bad.sum(), 1)) + 0.5 * init_jitter * rng.uniform(-1, 1, size=(bad.sum(), P))
        lp, g, _ = post.logp_dlogp(z, True)
    inv_mass = np.ones((chains, P))
    # dual averaging (Hoffman & Gelman 2014), per chain
    eps = np.full(chains, 0.1 / P ** 0.25)
    mu = np.log(10 * eps)
    hbar = np.zeros(chains)
    log_eps_bar = np.zeros(chains)
    gamma, t0, kappa = 0.05, 10.0, 0.75
    # running variance for the diagonal mass matrix (Welford), windowed
    wn = 0
    wmean = np.zeros((chains, P))
    wm2 = np.zeros((chains, P))
    win_end, win_len = 100, 100
    total = tune + draws
    out_z = np.empty((chains, draws, P))
    out_lp = np.empty((chains, draws))
    out_acc = np.empty((chains, draws))
    for it in range(total):
        tuning = it < tune
        mom = rng.standard_normal((chains, P)) / np.sqrt(inv_mass)
        h0 = -lp + 0.5 * np.sum(mom * mom * inv_mass, axis=1)
        nleap = int(rng.integers(max(1, max_leapfrog // 4), max_leapfrog + 1)) if path_length is None \
            else int(np.clip(np.ceil(path_length / np.median(eps)), 1, max_leapfrog))
        zn, gn, lpn, p = z.copy(), g.copy(), lp.copy(), mom.copy()
        alive = np.ones(chains, dtype=bool)
        for _ in range(nleap):
            p = p + 0.5 * eps[:, None] * gn
            zn = zn + eps[:, None] * inv_mass * p
            lpn, gn, _ = post.logp_dlogp(zn, True)
            dead = ~np.isfinite(lpn)
            if dead.any():                     # divergent chains: freeze them for the rest of the trajectory
                alive &= ~dead
                gn[dead] = 0.0
            p = p + 0.5 * eps[:, None] * gn
        with np.errstate(all='ignore'):
            h1 = -lpn + 0.5 * np.sum(p * p * inv_mass, axis=1)
        with np.errstate(over='ignore', invalid='ignore'):
            acc = np.where(alive & np.isfinite(h1), np.minimum(1.0, np.exp(h0 - h1)), 0.0)
        take = rng.uniform(size=chains) < acc
        z[take], g[take], lp[take] = zn[take], gn[take], lpn[take]
        if tuning:
            m = it + 1
            hbar = (1 - 1 / (m + t0)) * hbar + (target_accept - acc) / (m + t0)
            log_eps = mu - np.sqrt(m) / gamma * hbar
            eta = m ** (-kappa)
            log_eps_bar = eta * log_eps + (1 - eta) * log_eps_bar
            eps = np.exp(log_eps)
            wn += 1
            delta = z - wmean
            wmean += delta / wn
            wm2 += delta * (z - wmean)
            if m == win_end and m < tune - 50:
                var = wm2 / max(wn - 1, 1)
                inv_mass = (wn / (wn + 5.0)) * var + 1e-3 * (5.0 / (wn + 5.0))   # Stan's regularisation
                wn, wmean, wm2 = 0, np.zeros((chains, P)), np.zeros((chains, P))
                win_len *= 2
                win_end += win_len
                mu = np.log(10 * eps)
                hbar = np.zeros(chains)
            if m == tune:
                eps = np.exp(log_eps_bar)
        else:
            k = it - tune
            out_z[:, k], out_lp[:, k], out_acc[:, k] = z, lp, acc
        if progressbar and (it + 1) % 50 == 0:
            print(f'  iter {it + 1}/{total}  mean accept {acc.mean():.2f}  eps {np.median(eps):.3g}')
    # named posterior arrays
    posterior = {}
    theta = sp.theta_from_z(out_z)[0]
    for b, sl in zip(sp.blocks, sp.zslices):
        x = theta[..., b.theta_index]
        posterior[b.name] = x if (b.size > 1 or b.name in ('l', 'kv', 'iwgp', 'cwgp_pos', 'cwgp')) else x[..., 0]
        if b.transform is not None:
            zz = out_z[..., sl]
            posterior[b.tname] = zz if (b.size > 1 or b.name in ('l', 'kv', 'iwgp', 'cwgp_pos', 'cwgp')) else zz[..., 0]
    return Trace(posterior, out_lp, out_z, out_acc, eps)

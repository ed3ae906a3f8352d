"""MAP and MCMC drivers: the callers that turn batched logp/dlogp evaluations into ``fit`` results.

They stand where ``pm.find_MAP`` (andvaranaut/gpmcmc.py:332,345,357) and ``pm.sample`` (:351) stand in the
reference, and consume the same quantity -- log posterior and its gradient over the unconstrained variables --
but evaluated for MANY hyperparameter vectors per device call:
  * ``find_map``        SciPy L-BFGS-B exactly as PyMC drives it (objective -logp without transform Jacobian,
                        ``maxeval``, non-finite values mapped to a large penalty);
  * ``find_map_multi``  several independent L-BFGS-B runs (restarts) advanced in lock-step so every round of
                        objective evaluations is ONE batched call;
  * ``sample``          B chains of multinomial NUTS (what ``pm.sample`` runs, one process per chain there) advanced
                        together: every round is one leapfrog step of every chain = one batched call, and chains
                        move on to their next transition independently (continuous batching);
                        per-chain dual-averaging step size and windowed diagonal mass adaptation during tuning.
                        The sampled density (logp WITH Jacobian) is the reference's.
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
        if self.shard is None and hasattr(self.engine, 'loglik_grad_host_begin'):
            # the device evaluates (captured graph: upload, launches, one packed download) while the host forms the
            # Jacobian terms of the transforms and the priors: only theta itself is needed before the launch
            token = self.engine.loglik_grad_host_begin(self.space.theta_only(z))
            try:
                theta, dxdz, ljac, dljac = self.space.theta_from_z(z)
                lp, glp = self.space.prior(theta)
            finally:
                ll, gll, info = self.engine.loglik_grad_host_end(token)
            return self._assemble(z, ll, gll, info, lp, glp, dxdz, ljac, dljac, jacobian)
        theta, dxdz, ljac, dljac = self.space.theta_from_z(z)
        if self.shard is not None:
            ll, gll, info = self.shard.loglik_grad(self.engine, theta)
            lp, glp = self.space.prior(theta)
        elif hasattr(self.engine, 'loglik_grad_host'):
            ll, gll, info = self.engine.loglik_grad_host(theta)
            lp, glp = self.space.prior(theta)
        else:
            ll_t, g_t, info_t = self.engine.loglik_grad(theta)
            ll, gll, info = ll_t.cpu().numpy(), g_t.cpu().numpy(), info_t.cpu().numpy()
            if np.any(info < 0):
                raise RuntimeError('avn_gp_loglik_grad: factorisation aborted on the device (info = -1)')
            lp, glp = self.space.prior(theta)
        return self._assemble(z, ll, gll, info, lp, glp, dxdz, ljac, dljac, jacobian)

    def _assemble(self, z, ll, gll, info, lp, glp, dxdz, ljac, dljac, jacobian):
        self.n_eval += z.shape[0]
        self.n_calls += 1
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
        try:
            v, g, _ = self.post.logp_dlogp(z, jacobian=False)
            for k, i in enumerate(ids):
                self.results[i] = (v[k], g[k])
        except BaseException as e:   # device / collective failure: every waiting optimiser gets it, nobody is left waiting
            for i in ids:
                self.results[i] = e
        finally:
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
                    if not self.cv.wait(timeout=600.0):
                        raise RuntimeError('batched evaluation did not arrive within 600 s')
            r = self.results.pop(i)
            if isinstance(r, BaseException):
                raise r
            return r

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
    errors = []

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
        except (ValueError, FloatingPointError, np.linalg.LinAlgError) as e:
            print('Restart failed', e)       # a numerically failed restart must not block the others (gpmcmc.py:337-339)
        except BaseException as e:           # device errors end the whole fit: re-raised by the caller below
            errors.append(e)
        finally:
            rv.leave()
        out[i] = zi

    threads = [threading.Thread(target=run, args=(i,)) for i in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
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


def _popcount(n):
    c = np.zeros_like(n)
    while np.any(n > 0):
        c += n & 1
        n = n >> 1
    return c


def _trailing_ones(n):
    c = np.zeros_like(n)
    n = n.copy()
    while np.any(n & 1):
        m = (n & 1) == 1
        c += m
        n = np.where(m, n >> 1, n)
    return c


def _turning(inv_mass, p_a, p_b, p_sum):
    """generalised U-turn criterion (Betancourt 2017; PyMC / Stan): the summed momentum points backwards at either end."""
    with np.errstate(all='ignore'):
        return (np.sum(p_sum * inv_mass * p_a, axis=-1) <= 0.0) | (np.sum(p_sum * inv_mass * p_b, axis=-1) <= 0.0)


def _hmc_transition(post, rng, z, lp, g, eps, inv_mass, nleap):
    """one Metropolised trajectory of nleap leapfrog steps for every chain; returns the new state and the acceptance
    statistic the step-size adaptation consumes."""
    chains = z.shape[0]
    mom = rng.standard_normal(z.shape) / np.sqrt(inv_mass)
    h0 = -lp + 0.5 * np.sum(mom * mom * inv_mass, axis=1)
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
    z, g, lp = z.copy(), g.copy(), lp.copy()
    z[take], g[take], lp[take] = zn[take], gn[take], lpn[take]
    return z, lp, g, acc, np.full(chains, nleap), np.zeros(chains, dtype=bool)


class _NutsChains:
    """Multinomial NUTS for C chains whose transitions run ASYNCHRONOUSLY: every call of :meth:`step` is one leapfrog
    step of every chain that is inside a transition = one batched logp/dlogp call, and a chain whose tree has ended can
    start its next transition in the very next round instead of idling until the slowest tree of the batch is done
    (continuous batching; with lock-step transitions only ~35 % of the chain slots of a device call did work).

    The sampler is the one PyMC and Stan run (``pm.sample``, gpmcmc.py:351): the trajectory is doubled in a random
    direction until the generalised U-turn criterion fires, a leaf diverges (energy error above ``max_energy_error``)
    or ``max_treedepth`` is reached; the new state is drawn by progressive multinomial sampling (uniform inside a
    subtree, biased towards the new subtree at every doubling).  The recursion is unrolled: a subtree of 2^j leaves is
    built leaf by leaf and its internal U-turn checks use the O(j) checkpoint scheme (leaf index bit patterns tell which
    sub-subtrees end at this leaf), so the state per chain is fixed-size and chains need not be at the same depth --
    nor, now, in the same transition."""

    def __init__(self, post, rng, C, P, max_treedepth, max_energy_error=1000.0):
        self.post, self.rng, self.C, self.P = post, rng, C, P
        self.D, self.max_dE = max_treedepth, max_energy_error
        D = max_treedepth
        f = lambda *sh: np.zeros(sh)   # noqa: E731
        self.eps, self.inv_mass = f(C), np.ones((C, P))
        self.h0 = f(C)
        # main tree: both ends, proposal, log weight (energies relative to h0), momentum sum
        self.zl, self.pl, self.gl = f(C, P), f(C, P), f(C, P)
        self.zr, self.pr, self.gr = f(C, P), f(C, P), f(C, P)
        self.zp, self.lpp, self.gp = f(C, P), f(C), f(C, P)
        self.logw, self.psum = f(C), f(C, P)
        self.depth = np.zeros(C, dtype=np.int64)
        self.sum_acc = f(C)
        self.nprop = np.zeros(C, dtype=np.int64)
        self.diverged = np.zeros(C, dtype=bool)
        # subtree under construction
        self.going_right = np.zeros(C, dtype=bool)
        self.s_n = np.zeros(C, dtype=np.int64)               # leaves so far
        self.s_z, self.s_p, self.s_g = f(C, P), f(C, P), f(C, P)   # its growing edge (starts at the main tree's end)
        self.s_zp, self.s_lpp, self.s_gp = f(C, P), f(C), f(C, P)
        self.s_logw = np.full(C, -np.inf)
        self.s_psum = f(C, P)
        self.s_turn = np.zeros(C, dtype=bool)
        self.s_div = np.zeros(C, dtype=bool)
        self.ck_p = f(C, D + 1, P)
        self.ck_ps = f(C, D + 1, P)

    def begin(self, c, z, lp, g, eps, inv_mass):
        """start a transition of the chains with indices ``c`` from their current states (rows of z, lp, g)."""
        m = self.rng.standard_normal((len(c), self.P)) / np.sqrt(inv_mass)
        self.eps[c], self.inv_mass[c] = eps, inv_mass
        self.h0[c] = -lp + 0.5 * np.sum(m * m * inv_mass, axis=1)
        self.zl[c], self.pl[c], self.gl[c] = z, m, g
        self.zr[c], self.pr[c], self.gr[c] = z, m, g
        self.zp[c], self.lpp[c], self.gp[c] = z, lp, g
        self.logw[c] = 0.0
        self.psum[c] = m
        self.depth[c] = 0
        self.sum_acc[c] = 0.0
        self.nprop[c] = 0
        self.diverged[c] = False
        self._start_subtree(c)

    def _start_subtree(self, c):
        gr = self.rng.uniform(size=len(c)) < 0.5
        self.going_right[c] = gr
        r, l = c[gr], c[~gr]
        self.s_z[r], self.s_p[r], self.s_g[r] = self.zr[r], self.pr[r], self.gr[r]
        self.s_z[l], self.s_p[l], self.s_g[l] = self.zl[l], self.pl[l], self.gl[l]
        self.s_n[c] = 0
        self.s_logw[c] = -np.inf
        self.s_psum[c] = 0.0
        self.s_turn[c] = False
        self.s_div[c] = False

    def step(self, a):
        """one leapfrog step of the chains ``a`` (index array, all inside a transition): ONE batched device call.
        Returns the indices of the chains whose transition ended with this step; their results are read with
        :meth:`result`."""
        rng, inv_mass = self.rng, self.inv_mass
        v = np.where(self.going_right[a], 1.0, -1.0)[:, None] * self.eps[a, None]
        ph = self.s_p[a] + 0.5 * v * self.s_g[a]
        zn = self.s_z[a] + v * inv_mass[a] * ph
        lpn, gn, _ = self.post.logp_dlogp(zn, True)
        pn = ph + 0.5 * v * gn
        with np.errstate(all='ignore'):
            h1 = -lpn + 0.5 * np.sum(pn * pn * inv_mass[a], axis=1)
            dE = h1 - self.h0[a]
        dE = np.where(np.isfinite(dE), dE, np.inf)
        leaf_w = -dE
        div = dE > self.max_dE
        with np.errstate(over='ignore'):
            self.sum_acc[a] += np.minimum(1.0, np.exp(-dE))
        self.nprop[a] += 1
        self.s_z[a], self.s_p[a], self.s_g[a] = zn, pn, gn
        # uniform progressive sampling inside the subtree
        with np.errstate(all='ignore'):
            neww = np.logaddexp(self.s_logw[a], leaf_w)
            take = np.log(rng.uniform(size=len(a))) < leaf_w - neww
        take &= ~div
        ta = a[take]
        self.s_zp[ta], self.s_lpp[ta], self.s_gp[ta] = zn[take], lpn[take], gn[take]
        self.s_logw[a] = neww
        self.s_psum[a] += np.where(div[:, None], 0.0, pn)
        self.s_div[a] |= div
        # U-turn checks of the sub-subtrees that end at this leaf
        n = self.s_n[a]
        idx_max = _popcount(n >> 1)
        even = (n & 1) == 0
        ea = a[even]
        self.ck_p[ea, idx_max[even]] = pn[even]
        self.ck_ps[ea, idx_max[even]] = self.s_psum[ea]
        odd = ~even & ~div
        if odd.any():
            oa = a[odd]
            imax = idx_max[odd]
            imin = imax - _trailing_ones(n[odd]) + 1
            turn = np.zeros(len(oa), dtype=bool)
            for k in range(int(imax.max()), int(imin.min()) - 1, -1):
                m = (k <= imax) & (k >= imin) & ~turn
                if not m.any():
                    continue
                om = oa[m]
                sub = self.s_psum[om] - self.ck_ps[om, k] + self.ck_p[om, k]
                turn[m] = _turning(inv_mass[om], self.ck_p[om, k], pn[odd][m], sub)
            self.s_turn[oa] |= turn
        self.s_n[a] += 1
        # finished subtrees are merged into the main tree
        done = (self.s_n[a] == (1 << self.depth[a])) | self.s_turn[a] | self.s_div[a]
        if not done.any():
            return a[:0]
        da = a[done]
        ok = ~(self.s_turn[da] | self.s_div[da])
        with np.errstate(all='ignore'):
            tp = np.where(ok, np.exp(np.minimum(0.0, self.s_logw[da] - self.logw[da])), 0.0)
        mv = da[rng.uniform(size=len(da)) < tp]
        self.zp[mv], self.lpp[mv], self.gp[mv] = self.s_zp[mv], self.s_lpp[mv], self.s_gp[mv]
        r = da[self.going_right[da]]
        self.zr[r], self.pr[r], self.gr[r] = self.s_z[r], self.s_p[r], self.s_g[r]
        l = da[~self.going_right[da]]
        self.zl[l], self.pl[l], self.gl[l] = self.s_z[l], self.s_p[l], self.s_g[l]
        self.psum[da] += self.s_psum[da]
        self.logw[da] = np.logaddexp(self.logw[da], self.s_logw[da])
        self.depth[da] += 1
        self.diverged[da] |= self.s_div[da]
        stop = ~ok | _turning(inv_mass[da], self.pl[da], self.pr[da], self.psum[da]) | (self.depth[da] >= self.D)
        cont = da[~stop]
        if len(cont):
            self._start_subtree(cont)
        return da[stop]

    def result(self, c):
        """(z, lp, g, mean leaf acceptance, leapfrog steps, diverged) of the finished transitions of chains ``c``"""
        return (self.zp[c].copy(), self.lpp[c].copy(), self.gp[c].copy(), self.sum_acc[c] / np.maximum(self.nprop[c], 1),
                self.nprop[c].copy(), self.diverged[c].copy())


def _nuts_transition(post, rng, z, lp, g, eps, inv_mass, max_treedepth, max_energy_error=1000.0):
    """One No-U-Turn transition for every chain with the chains advanced in LOCK-STEP (all start together, finished
    trees wait for the slowest): the building block of :class:`_NutsChains` driven synchronously; ``sample`` itself
    runs the chains asynchronously."""
    C, P = z.shape
    nc = _NutsChains(post, rng, C, P, max_treedepth, max_energy_error)
    allc = np.arange(C)
    nc.begin(allc, z, lp, g, eps, inv_mass)
    active = np.ones(C, dtype=bool)
    while active.any():
        active[nc.step(np.where(active)[0])] = False
    return nc.result(allc)


def sample(post, draws=1000, tune=1000, chains=4, seed=None, target_accept=0.8, path_length=None,
           max_leapfrog=64, start_z=None, init_jitter=1.0, progressbar=False, sampler='nuts', max_treedepth=10):
    """``chains`` Markov chains advanced together; every leapfrog step is one batched logp/dlogp call.
    ``sampler='nuts'`` (default, what ``pm.sample`` runs): multinomial NUTS with asynchronous transitions (continuous
    batching: a finished tree starts its next transition at once), see :class:`_NutsChains`;
    ``sampler='hmc'``: fixed-length trajectories (jittered length, at most ``max_leapfrog`` steps).  Both share PyMC's
    ``jitter+adapt_diag`` start, dual-averaging step-size adaptation towards ``target_accept`` and a windowed
    diagonal mass-matrix estimate during tuning."""
    if sampler not in ('nuts', 'hmc'):
        raise ValueError("sampler must be 'nuts' or 'hmc'")
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
    # running variance for the diagonal mass matrix (Welford), windowed; every chain keeps its own iteration count
    wn = np.zeros(chains)
    wmean = np.zeros((chains, P))
    wm2 = np.zeros((chains, P))
    win_end = np.full(chains, 100)
    win_len = np.full(chains, 100)
    total = tune + draws
    out_z = np.empty((chains, draws, P))
    out_lp = np.empty((chains, draws))
    out_acc = np.empty((chains, draws))
    out_n = np.empty((chains, draws), dtype=np.int64)
    out_div = np.zeros((chains, draws), dtype=bool)
    it = np.zeros(chains, dtype=np.int64)        # transitions finished so far, per chain

    def finish(c, zc, lpc, gc, acc, nst, dv):
        """transition results of the chains ``c`` (index array): adaptation while tuning, else record the draw"""
        nonlocal eps, inv_mass, mu, hbar, log_eps_bar
        z[c], lp[c], g[c] = zc, lpc, gc
        m = it[c] + 1
        tun = m <= tune
        ct, mt = c[tun], m[tun].astype(np.float64)
        if len(ct):
            hbar[ct] = (1 - 1 / (mt + t0)) * hbar[ct] + (target_accept - acc[tun]) / (mt + t0)
            log_eps = mu[ct] - np.sqrt(mt) / gamma * hbar[ct]
            eta = mt ** (-kappa)
            log_eps_bar[ct] = eta * log_eps + (1 - eta) * log_eps_bar[ct]
            eps[ct] = np.exp(log_eps)
            wn[ct] += 1
            delta = z[ct] - wmean[ct]
            wmean[ct] += delta / wn[ct, None]
            wm2[ct] += delta * (z[ct] - wmean[ct])
            w = (m[tun] == win_end[ct]) & (m[tun] < tune - 50)
            cw = ct[w]
            if len(cw):
                var = wm2[cw] / np.maximum(wn[cw, None] - 1, 1)
                inv_mass[cw] = (wn[cw, None] / (wn[cw, None] + 5.0)) * var + 1e-3 * (5.0 / (wn[cw, None] + 5.0))   # Stan's regularisation
                wn[cw], wmean[cw], wm2[cw] = 0, 0.0, 0.0
                win_len[cw] *= 2
                win_end[cw] += win_len[cw]
                mu[cw] = np.log(10 * eps[cw])
                hbar[cw] = 0.0
            last = ct[m[tun] == tune]
            eps[last] = np.exp(log_eps_bar[last])
        cd = c[~tun]
        if len(cd):
            k = m[~tun] - 1 - tune
            out_z[cd, k], out_lp[cd, k], out_acc[cd, k] = zc[~tun], lpc[~tun], acc[~tun]
            out_n[cd, k], out_div[cd, k] = nst[~tun], dv[~tun]
        it[c] += 1

    allc = np.arange(chains)
    if sampler == 'nuts':
        # continuous batching: one round = one leapfrog of EVERY chain that still has transitions to do; a chain whose
        # tree ended starts its next transition in the next round (its adaptation state is its own)
        nuts = _NutsChains(post, rng, chains, P, max_treedepth)
        if total > 0:
            nuts.begin(allc, z, lp, g, eps, inv_mass)
        running = np.full(chains, total > 0)
        rounds = 0
        while running.any():
            fin = nuts.step(np.where(running)[0])
            rounds += 1
            if len(fin):
                finish(fin, *nuts.result(fin))
                again = fin[it[fin] < total]
                running[fin[it[fin] >= total]] = False
                if len(again):
                    nuts.begin(again, z[again], lp[again], g[again], eps[again], inv_mass[again])
            if progressbar and rounds % 200 == 0:
                print(f'  round {rounds}: transitions done min {it.min()} / mean {it.mean():.0f} of {total}')
    else:
        for _ in range(total):
            nleap = int(rng.integers(max(1, max_leapfrog // 4), max_leapfrog + 1)) if path_length is None \
                else int(np.clip(np.ceil(path_length / np.median(eps)), 1, max_leapfrog))
            finish(allc, *_hmc_transition(post, rng, z, lp, g, eps, inv_mass, nleap))
            if progressbar and it[0] % 50 == 0:
                print(f'  iter {it[0]}/{total}  eps {np.median(eps):.3g}')
    # named posterior arrays
    posterior = {}
    theta = sp.theta_from_z(out_z)[0]
    for b, sl in zip(sp.blocks, sp.zslices):
        x = theta[..., b.theta_index]
        posterior[b.name] = x if (b.size > 1 or b.name in sp.VECTOR_NAMES) else x[..., 0]
        if b.transform is not None:
            zz = out_z[..., sl]
            posterior[b.tname] = zz if (b.size > 1 or b.name in sp.VECTOR_NAMES) else zz[..., 0]
    tr = Trace(posterior, out_lp, out_z, out_acc, eps)
    tr.sample_stats['n_steps'] = out_n
    tr.sample_stats['diverging'] = out_div
    tr.sample_stats['tree_depth'] = np.ceil(np.log2(np.maximum(out_n, 1) + 1)).astype(np.int64)
    return tr

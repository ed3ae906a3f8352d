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
        elif hasattr(self.engine, 'loglik_grad_host'):
            ll, gll, info = self.engine.loglik_grad_host(theta)      # one packed device->host copy
        else:
            ll_t, g_t, info_t = self.engine.loglik_grad(theta)
            ll, gll, info = ll_t.cpu().numpy(), g_t.cpu().numpy(), info_t.cpu().numpy()
            if np.any(info < 0):
                raise RuntimeError('avn_gp_loglik_grad: factorisation aborted on the device (info = -1)')
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


def _nuts_transition(post, rng, z, lp, g, eps, inv_mass, max_treedepth, max_energy_error=1000.0):
    """One No-U-Turn transition for every chain, the chains advanced in LOCK-STEP: each round of the loop below is one
    leapfrog step of every chain that is still growing its trajectory = one batched logp/dlogp call.

    Multinomial NUTS as PyMC and Stan run it (the sampler behind ``pm.sample``, gpmcmc.py:351): the trajectory is
    doubled in a random direction until the generalised U-turn criterion fires, a leaf diverges (energy error above
    ``max_energy_error``) or ``max_treedepth`` is reached; the new state is drawn by progressive multinomial sampling
    (uniform inside a subtree, biased towards the new subtree at every doubling).  The recursion is unrolled: a
    subtree of 2^j leaves is built leaf by leaf and its internal U-turn checks use the O(j) checkpoint scheme
    (leaf index bit patterns tell which sub-subtrees end at this leaf), so the state per chain is fixed-size and the
    chains need not be at the same depth."""
    C, P = z.shape
    D = max_treedepth
    mom = rng.standard_normal(z.shape) / np.sqrt(inv_mass)
    h0 = -lp + 0.5 * np.sum(mom * mom * inv_mass, axis=1)
    # main tree: both ends, proposal, log weight (energies relative to h0), momentum sum
    zl, pl, gl = z.copy(), mom.copy(), g.copy()
    zr, pr, gr = z.copy(), mom.copy(), g.copy()
    zp, lpp, gp = z.copy(), lp.copy(), g.copy()
    logw = np.zeros(C)
    psum = mom.copy()
    depth = np.zeros(C, dtype=np.int64)
    sum_acc = np.zeros(C)
    nprop = np.zeros(C, dtype=np.int64)
    diverged = np.zeros(C, dtype=bool)
    active = np.ones(C, dtype=bool)
    # subtree under construction
    going_right = rng.uniform(size=C) < 0.5
    s_n = np.zeros(C, dtype=np.int64)               # leaves so far
    s_z, s_p, s_g = z.copy(), mom.copy(), g.copy()  # its growing edge (starts at the main tree's end)
    s_zp, s_lpp, s_gp = z.copy(), lp.copy(), g.copy()
    s_logw = np.full(C, -np.inf)
    s_psum = np.zeros((C, P))
    s_turn = np.zeros(C, dtype=bool)
    s_div = np.zeros(C, dtype=bool)
    s_pfirst = np.zeros((C, P))
    ck_p = np.zeros((C, D + 1, P))
    ck_ps = np.zeros((C, D + 1, P))

    def start_subtree(m):
        going_right[m] = rng.uniform(size=int(m.sum())) < 0.5
        right = m & going_right
        left = m & ~going_right
        s_z[right], s_p[right], s_g[right] = zr[right], pr[right], gr[right]
        s_z[left], s_p[left], s_g[left] = zl[left], pl[left], gl[left]
        s_n[m] = 0
        s_logw[m] = -np.inf
        s_psum[m] = 0.0
        s_turn[m] = False
        s_div[m] = False

    start_subtree(active.copy())
    while active.any():
        a = np.where(active)[0]
        v = np.where(going_right[a], 1.0, -1.0)[:, None] * eps[a, None]
        ph = s_p[a] + 0.5 * v * s_g[a]
        zn = s_z[a] + v * inv_mass[a] * ph
        lpn, gn, _ = post.logp_dlogp(zn, True)
        pn = ph + 0.5 * v * gn
        with np.errstate(all='ignore'):
            h1 = -lpn + 0.5 * np.sum(pn * pn * inv_mass[a], axis=1)
            dE = h1 - h0[a]
        dE = np.where(np.isfinite(dE), dE, np.inf)
        leaf_w = -dE
        div = dE > max_energy_error
        with np.errstate(over='ignore'):
            sum_acc[a] += np.minimum(1.0, np.exp(-dE))
        nprop[a] += 1
        s_z[a], s_p[a], s_g[a] = zn, pn, gn
        first = s_n[a] == 0
        s_pfirst[a[first]] = pn[first]
        # uniform progressive sampling inside the subtree
        with np.errstate(all='ignore'):
            neww = np.logaddexp(s_logw[a], leaf_w)
            take = np.log(rng.uniform(size=len(a))) < leaf_w - neww
        take &= ~div
        ta = a[take]
        s_zp[ta], s_lpp[ta], s_gp[ta] = zn[take], lpn[take], gn[take]
        s_logw[a] = neww
        s_psum[a] += np.where(div[:, None], 0.0, pn)
        s_div[a] |= div
        # U-turn checks of the sub-subtrees that end at this leaf
        n = s_n[a]
        idx_max = _popcount(n >> 1)
        even = (n & 1) == 0
        ea = a[even]
        ck_p[ea, idx_max[even]] = pn[even]
        ck_ps[ea, idx_max[even]] = s_psum[ea]
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
                sub = s_psum[om] - ck_ps[om, k] + ck_p[om, k]
                turn[m] = _turning(inv_mass[om], ck_p[om, k], pn[odd][m], sub)
            s_turn[oa] |= turn
        s_n[a] += 1
        # finished subtrees are merged into the main tree
        done = (s_n[a] == (1 << depth[a])) | s_turn[a] | s_div[a]
        if not done.any():
            continue
        da = a[done]
        ok = ~(s_turn[da] | s_div[da])
        with np.errstate(all='ignore'):
            tp = np.where(ok, np.exp(np.minimum(0.0, s_logw[da] - logw[da])), 0.0)
        mv = da[rng.uniform(size=len(da)) < tp]
        zp[mv], lpp[mv], gp[mv] = s_zp[mv], s_lpp[mv], s_gp[mv]
        r = da[going_right[da]]
        zr[r], pr[r], gr[r] = s_z[r], s_p[r], s_g[r]
        l = da[~going_right[da]]
        zl[l], pl[l], gl[l] = s_z[l], s_p[l], s_g[l]
        psum[da] += s_psum[da]
        logw[da] = np.logaddexp(logw[da], s_logw[da])
        depth[da] += 1
        diverged[da] |= s_div[da]
        stop = ~ok | _turning(inv_mass[da], pl[da], pr[da], psum[da]) | (depth[da] >= D)
        active[da[stop]] = False
        cont = np.zeros(C, dtype=bool)
        cont[da[~stop]] = True
        if cont.any():
            start_subtree(cont)
    return zp, lpp, gp, sum_acc / np.maximum(nprop, 1), nprop, diverged


def sample(post, draws=1000, tune=1000, chains=4, seed=None, target_accept=0.8, path_length=None,
           max_leapfrog=64, start_z=None, init_jitter=1.0, progressbar=False, sampler='nuts', max_treedepth=10):
    """``chains`` Markov chains advanced in lock-step; every leapfrog step is one batched logp/dlogp call.
    ``sampler='nuts'`` (default, what ``pm.sample`` runs): multinomial NUTS, see :func:`_nuts_transition`;
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
    # running variance for the diagonal mass matrix (Welford), windowed
    wn = 0
    wmean = np.zeros((chains, P))
    wm2 = np.zeros((chains, P))
    win_end, win_len = 100, 100
    total = tune + draws
    out_z = np.empty((chains, draws, P))
    out_lp = np.empty((chains, draws))
    out_acc = np.empty((chains, draws))
    out_n = np.empty((chains, draws), dtype=np.int64)
    out_div = np.zeros((chains, draws), dtype=bool)
    for it in range(total):
        tuning = it < tune
        if sampler == 'nuts':
            z, lp, g, acc, nst, dv = _nuts_transition(post, rng, z, lp, g, eps, inv_mass, max_treedepth)
        else:
            nleap = int(rng.integers(max(1, max_leapfrog // 4), max_leapfrog + 1)) if path_length is None \
                else int(np.clip(np.ceil(path_length / np.median(eps)), 1, max_leapfrog))
            z, lp, g, acc, nst, dv = _hmc_transition(post, rng, z, lp, g, eps, inv_mass, nleap)
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
            out_z[:, k], out_lp[:, k], out_acc[:, k], out_n[:, k], out_div[:, k] = z, lp, acc, nst, dv
        if progressbar and (it + 1) % 50 == 0:
            print(f'  iter {it + 1}/{total}  mean accept {acc.mean():.2f}  eps {np.median(eps):.3g}')
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

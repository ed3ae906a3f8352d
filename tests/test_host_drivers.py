"""CPU tests of the host logic above the C ABI: parameter space / priors / transforms against scipy and finite
differences, MAP and lock-step HMC drivers against the oracle (through a test double of the engine), and the
world_size-2 sharding path on gloo."""
import os
import sys

import numpy as np
import pytest
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import gp_oracle as go  # noqa: E402
from andvaranaut_b200.priors import ParamSpace  # noqa: E402
from andvaranaut_b200 import drivers  # noqa: E402
import cases  # noqa: E402
from fake_engine import OracleEngine  # noqa: E402


def test_param_space_layout_and_names():
    sp = ParamSpace(nx=3, nkern=2, noise=True, n_iw=4, cw_pos=[False, True, False, True], has_alpha=True)
    assert sp.P == 1 + 6 + 2 + 4 + 4 + 1
    assert [b.name for b in sp.blocks] == ['gv', 'l', 'kv', 'iwgp', 'cwgp_pos', 'cwgp', 'alpha']
    z = sp.initial_z()
    h = sp.hypers_dict(z)
    # PyMC moments: HalfNormal sigma, LogNormal exp(mu + sigma^2/2), Normal mu
    assert np.isclose(h['gv'], 1e-3) and np.allclose(h['l'], np.exp(0.5)) and np.allclose(h['kv'], np.exp(0.56 + 0.75 ** 2 / 2))
    assert np.allclose(h['iwgp'], np.exp(0.25 ** 2 / 2)) and np.allclose(h['cwgp'], 0.0)
    assert set(h) >= {'gv_log__', 'l_log__', 'kv_log__', 'iwgp_log__', 'cwgp_pos_log__', 'cwgp', 'alpha_log__', 'alpha'}
    th = sp.theta_from_z(z)[0]
    # cw entries interleave in wgp order: free, pos, free, pos
    assert np.allclose(th[13:17], [0.0, np.exp(0.25 ** 2 / 2), 0.0, np.exp(0.25 ** 2 / 2)])
    assert np.allclose(sp.z_from_theta(th), z)
    assert np.allclose(sp.theta_from_hypers(h), th)
    # start dict overrides, constrained or transformed names
    z2 = sp.initial_z({'l': np.full(6, 2.0), 'kv_log__': np.array([0.1, 0.2])})
    h2 = sp.hypers_dict(z2)
    assert np.allclose(h2['l'], 2.0) and np.allclose(h2['kv_log__'], [0.1, 0.2])


@pytest.mark.parametrize('truncate', [False, True])
def test_prior_logp_and_gradients(truncate):
    sp = ParamSpace(nx=2, nkern=1, noise=True, n_iw=2, cw_pos=[False, True], truncate=truncate)
    rng = np.random.default_rng(0)
    z = sp.initial_z() + 0.1 * rng.normal(size=sp.P)
    theta, dxdz, ljac, dljac = sp.theta_from_z(z)
    lp, g = sp.prior(theta)
    # against scipy.stats
    if truncate:
        ref = st.truncnorm((1e-15 - 0) / 1e-3, (1 - 0) / 1e-3, loc=0, scale=1e-3).logpdf(theta[0])
        ref += st.truncnorm((1e-3 - .5) / .15, (100 - .5) / .15, loc=.5, scale=.15).logpdf(theta[1:3]).sum()
        ref += st.truncnorm((.1 - 1) / .15, (100 - 1) / .15, loc=1, scale=.15).logpdf(theta[3])
        ref += st.truncnorm((1e-3 - 1) / 1, (5 - 1) / 1, loc=1, scale=1).logpdf(theta[4:6]).sum()
        ref += st.truncnorm(-10, 10).logpdf(theta[6]) + st.truncnorm((1e-3 - 1), 4, loc=1, scale=1).logpdf(theta[7])
    else:
        ref = st.halfnorm(scale=1e-3).logpdf(theta[0]) + st.lognorm(s=1).logpdf(theta[1:3]).sum()
        ref += st.lognorm(s=0.75, scale=np.exp(0.56)).logpdf(theta[3]) + st.lognorm(s=0.25).logpdf(theta[4:6]).sum()
        ref += st.norm().logpdf(theta[6]) + st.lognorm(s=0.25).logpdf(theta[7])
    assert np.isclose(lp, ref, rtol=1e-12)
    # d/dz of (prior + log-Jacobian) by central differences
    def f(zz):
        t, _, lj, _ = sp.theta_from_z(zz)
        return sp.prior(t)[0] + lj
    gz = sp.grad_theta_to_z(g, dxdz) + dljac
    h = 1e-5
    noise = 1e-15 * abs(f(z)) / h          # rounding floor of the difference quotient
    for i in range(sp.P):
        e = np.zeros(sp.P)
        e[i] = h
        assert np.isclose((f(z + e) - f(z - e)) / (2 * h), gz[i], rtol=1e-5, atol=1e-5 + 10 * noise)


def small_problem(seed=3, N=40):
    spec = go.ModelSpec(nx=2, kerns=['Matern52'], noise=True)
    X, y, th, _ = cases.synth(spec, N, seed=seed)
    eng = OracleEngine(spec)
    eng.set_data(X, y)
    sp = ParamSpace(nx=2, nkern=1, noise=True)
    return spec, X, y, eng, sp


def test_posterior_gradient_matches_finite_differences():
    spec, X, y, eng, sp = small_problem()
    post = drivers.Posterior(eng, sp)
    z = sp.initial_z() + 0.05
    for jac in (False, True):
        v, g, info = post.logp_dlogp(z[None, :], jac)
        for i in range(sp.P):
            e = np.zeros(sp.P)
            e[i] = 1e-5
            fd = (post.logp_dlogp((z + e)[None, :], jac)[0][0] - post.logp_dlogp((z - e)[None, :], jac)[0][0]) / 2e-5
            assert np.isclose(fd, g[0, i], rtol=1e-5, atol=1e-5)


def test_find_map_reaches_stationary_point_and_counts_evals():
    spec, X, y, eng, sp = small_problem()
    post = drivers.Posterior(eng, sp)
    z, lp, nev = drivers.find_map(post, sp.initial_z())
    v, g, _ = post.logp_dlogp(z[None, :], False)
    assert np.max(np.abs(g)) < 1e-2 * max(1.0, abs(lp)) and lp > post.logp_dlogp(sp.initial_z()[None, :], False)[0][0]
    assert 5 < nev < 200          # tutorial band: 11-39 evaluations at N ~ 100


def test_batched_restarts_share_device_calls_and_agree_with_sequential():
    spec, X, y, eng, sp = small_problem()
    post = drivers.Posterior(eng, sp)
    rng = np.random.default_rng(1)
    z0s = np.vstack([sp.initial_z(), sp.initial_z() + 0.3 * rng.normal(size=(3, sp.P))])
    eng.calls.clear()
    zs, lps = drivers.find_map_multi(post, z0s)
    assert max(eng.calls) == 4 and len(eng.calls) < 200     # evaluations really are batched
    for i in range(4):
        zi, lpi, _ = drivers.find_map(drivers.Posterior(eng, sp), z0s[i])
        assert np.isclose(lps[i], lpi, rtol=1e-6, atol=1e-4)


def test_hmc_samples_the_posterior():
    """statistical check on a 3-parameter model: chain means of log-hyperparameters agree with a long
    random-walk Metropolis run on the same oracle density."""
    spec, X, y, eng, sp = small_problem(N=25)
    post = drivers.Posterior(eng, sp)
    tr = drivers.sample(post, draws=150, tune=150, chains=6, seed=5, max_leapfrog=12)
    z = tr.z.reshape(-1, sp.P)
    assert np.isfinite(tr.sample_stats['lp']).all() and tr.sample_stats['acceptance_rate'].mean() > 0.5
    assert tr.posterior['l'].shape == (6, 150, 2) and tr.posterior['gv'].shape == (6, 150)
    # Metropolis reference
    rng = np.random.default_rng(9)
    zc = drivers.find_map(post, sp.initial_z())[0]
    lpc = post.logp_dlogp(zc[None, :], True)[0][0]
    acc = []
    step = 0.35 * z.std(axis=0)
    for it in range(4000):
        prop = zc + step * rng.normal(size=sp.P)
        lpp = post.logp_dlogp(prop[None, :], True)[0][0]
        if np.log(rng.uniform()) < lpp - lpc:
            zc, lpc = prop, lpp
        if it >= 500:
            acc.append(zc.copy())
    ref = np.array(acc)
    tol = 0.35 * ref.std(axis=0) + 0.05
    assert np.all(np.abs(z.mean(axis=0) - ref.mean(axis=0)) < tol), (z.mean(axis=0), ref.mean(axis=0), tol)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from andvaranaut_b200.dist import Shard
    spec, X, y, eng, sp = small_problem()
    rng = np.random.default_rng(0)
    thetas = sp.theta_from_z(sp.initial_z() + 0.1 * rng.normal(size=(5, sp.P)))[0]   # 5 samples over 2 ranks: ragged
    sh = Shard()
    ll, g, info = sh.loglik_grad(eng, thetas)
    eng.factorize(thetas[0])
    Xs = rng.uniform(0, 1, (7, 2))
    mu, var = sh.predict(eng, Xs)
    q.put((rank, ll, g, info, mu, var, list(eng.calls)))
    dist.destroy_process_group()


def test_sharding_world_size_2_gloo_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    spec, X, y, eng, sp = small_problem()
    rng = np.random.default_rng(0)
    thetas = sp.theta_from_z(sp.initial_z() + 0.1 * rng.normal(size=(5, sp.P)))[0]
    ll, g, info = eng.loglik_grad(thetas)
    eng.factorize(thetas[0])
    mu, var = eng.predict(rng.uniform(0, 1, (7, 2)))
    for r in res:
        # gathered results are bit-identical to the unsharded evaluation on every rank
        assert np.array_equal(r[1], ll.numpy()) and np.array_equal(r[2], g.numpy()) and np.array_equal(r[3], info.numpy())
        assert np.array_equal(r[4], mu.numpy()) and np.array_equal(r[5], var.numpy())
    assert res[0][6][0] == 3 and res[1][6][0] == 2      # 5 samples split 3 + 2


# ---- x models: inverse problem, BO 'map' acquisition model, NUTS ------------------------------------------------
def test_xspace_follows_the_reference_prior_conversion():
    """gpmcmc.py:1053-1096: uniform -> Uniform(support), norm -> Normal(mean, std), truncnorm -> TruncatedNormal with
    (mu, sigma) recovered from args / kwds in every calling convention the reference distinguishes."""
    from andvaranaut_b200.priors import XSpace
    tn = [st.truncnorm(-1.0, 2.0, 0.5, 0.2), st.truncnorm(-1.0, 2.0, 0.5, scale=0.2), st.truncnorm(-1.0, 2.0, 0.5),
          st.truncnorm(a=-1.0, b=2.0, loc=0.5, scale=0.2), st.truncnorm(-1.0, 2.0), st.truncnorm(-1.0, b=2.0, scale=0.2)]
    expect = [(0.5, 0.2), (0.5, 0.2), (0.5, 1.0), (0.5, 0.2), (0.0, 1.0), (0.0, 0.2)]
    sp = XSpace([st.uniform(1.0, 0.5), st.norm(2.0, 3.0)] + tn)
    assert [b.name for b in sp.blocks] == [f'x{k}' for k in range(8)]
    assert sp.blocks[0].family == 'uniform' and sp.blocks[0].args[2:] == (1.0, 1.5)
    assert sp.blocks[1].family == 'normal' and sp.blocks[1].args == (2.0, 3.0)
    for b, pr, (mu, sg) in zip(sp.blocks[2:], tn, expect):
        assert b.family == 'truncnormal' and np.allclose(b.args, (mu, sg, mu - sg, mu + 2 * sg))
    # densities against scipy, names as PyMC reports them
    rng = np.random.default_rng(1)
    z = rng.normal(size=sp.P)
    x, dxdz, ljac, dljac = sp.theta_from_z(z)
    lp, _ = sp.prior(x)
    ref = st.uniform(1.0, 0.5).logpdf(x[0]) + st.norm(2.0, 3.0).logpdf(x[1]) + sum(p.logpdf(v) for p, v in zip(tn, x[2:]))
    assert np.isclose(lp, ref, rtol=1e-12)
    h = sp.hypers_dict(z)
    assert set(h) == {f'x{k}' for k in range(8)} | {'x0_interval__'} | {f'x{k}_interval__' for k in range(2, 8)}
    assert h['x0'].shape == ()
    with pytest.raises(Exception, match='not implemented'):
        XSpace([st.beta(2, 3)])


def _inverse_case(kerns, ops, nobs, with_var):
    rng = np.random.default_rng(8)
    N, d = 40, 2
    X = rng.uniform(0.05, 0.95, size=(N, d))
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.01 * rng.normal(size=N)
    nk = len(kerns)
    hyp = dict(gv=np.array(2e-4), l=rng.uniform(0.4, 0.9, d * nk), kv=rng.uniform(0.8, 1.5, nk))
    if 'RatQuad' in kerns:
        hyp['alpha'] = np.array(1.7)
    yo = np.full(nobs, 0.9) + 0.01 * np.arange(nobs)
    noise_o = 3e-3 if with_var else 0.0
    return X, y, hyp, yo, noise_o


@pytest.mark.parametrize('kerns,ops,nobs,with_var', [(['RBF'], [], 1, False), (['Matern52'], [], 1, True),
                                                    (['Matern32', 'RBF'], ['+'], 3, True),
                                                    (['Exponential', 'RatQuad'], ['*'], 2, True)])
def test_inverse_likelihood_block_form_equals_dense_reference_graph(kerns, ops, nobs, with_var):
    """The block (Schur) evaluation of the inverse-problem potential == the oracle's literal restatement of
    gpmcmc.py:1098-1165 (dense Cholesky of the stacked system), incl. the sqrt-variance-on-the-diagonal quirk and the
    full-form kernel diagonal; gradient against central differences.  Engine = oracle test double (CPU)."""
    from andvaranaut_b200.xpost import InverseLikelihood, kdiag_values
    X, y, hyp, yo, noise_o = _inverse_case(kerns, ops, nobs, with_var)
    jitter = 1e-6
    d = X.shape[1]
    noise_t = np.sqrt(hyp['gv'] + jitter)
    spec = go.ModelSpec(nx=d, kerns=kerns, ops=ops, noise=True, jitter=0.0)
    sp = ParamSpace(d, len(kerns), True, has_alpha='RatQuad' in kerns)
    th = sp.theta_from_hypers(dict(hyp, gv=noise_t))
    eng = OracleEngine(spec)
    eng.set_data(X, y)
    eng.factorize(th)
    const = go.loglik(spec, th, X, y, want_grad=False).ll
    cfull, cdiag = kdiag_values(kerns, ops, hyp['kv'], float(hyp.get('alpha', 1.0)))
    # a conversion with a non-trivial derivative: xc = x^2 on [0,1]
    pot = InverseLikelihood(eng, lambda x: (x ** 2, 2 * x), yo, noise_o, cfull - cdiag, const)
    xq = np.random.default_rng(2).uniform(0.2, 0.9, size=(5, d))
    val, gx = pot(xq)
    ynoise = np.r_[np.full(len(y), noise_t), np.full(nobs, noise_o)]
    th_ref = sp.theta_from_hypers(hyp)

    def dense(x):
        return go.inverse_loglik(spec, th_ref, X, y, x ** 2, yo, ynoise)
    ref = np.array([dense(x) for x in xq])
    assert np.max(np.abs(val - ref)) <= 1e-9 * np.max(np.abs(ref))
    for i in range(d):
        h = np.zeros(d)
        h[i] = 1e-6
        fd = np.array([(dense(x + h) - dense(x - h)) / 2e-6 for x in xq])
        # the Exponential kernel's cusp at training points makes the central difference itself only ~1e-5 accurate
        tol = 1e-4 if 'Exponential' in kerns else 1e-5
        assert np.max(np.abs(gx[:, i] - fd)) <= tol * max(1.0, np.max(np.abs(fd)))


def test_inverse_map_and_nuts_recover_the_point():
    """find_map / NUTS over the x model: a 1-observation inverse problem on a monotone response has its posterior
    mass where the surrogate reproduces the observation."""
    from andvaranaut_b200.xpost import InverseLikelihood, XPosterior, kdiag_values
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 1, size=(30, 1))
    y = 2.0 * X[:, 0] + 0.3 * X[:, 0] ** 2
    hyp = dict(gv=np.array(1e-5), l=np.array([1.2]), kv=np.array([3.0]))
    spec = go.ModelSpec(nx=1, kerns=['RBF'], noise=True, jitter=0.0)
    sp = ParamSpace(1, 1, True)
    th = sp.theta_from_hypers(dict(hyp, gv=np.sqrt(hyp['gv'] + 1e-6)))
    eng = OracleEngine(spec)
    eng.set_data(X, y)
    eng.factorize(th)
    xtrue = 0.62
    yo = np.array([2.0 * xtrue + 0.3 * xtrue ** 2])
    pot = InverseLikelihood(eng, lambda x: (x, np.ones_like(x)), yo, 0.0, 0.0, 0.0)
    post = XPosterior([st.uniform(0, 1)], pot)
    zs, lps = drivers.find_map_multi(post, rng.standard_normal((4, 1)))
    xs = post.space.theta_from_z(zs)[0][:, 0]
    assert np.all(np.abs(xs - xtrue) < 0.02), xs
    tr = drivers.sample(post, draws=120, tune=150, chains=6, seed=4)
    xm = tr.posterior['x0']
    assert xm.shape == (6, 120) and abs(xm.mean() - xtrue) < 0.03
    assert 'x0_interval__' in tr.posterior and tr.sample_stats['diverging'].mean() < 0.05


def test_nuts_and_hmc_sample_a_known_density():
    """lock-step NUTS (default) and HMC on a correlated Gaussian x a uniform: means, variances, covariance."""
    from andvaranaut_b200.xpost import XPosterior

    def pot(x):
        dlt = x[:, 0] - x[:, 1]
        g = np.zeros_like(x)
        g[:, 0], g[:, 1] = -dlt, dlt
        return -0.5 * dlt * dlt, g
    lam = np.array([[1.25, -1.0], [-1.0, 5.0]])
    cov = np.linalg.inv(lam)
    mean = cov @ np.array([0.25, -4.0])
    for smp in ('nuts', 'hmc'):
        post = XPosterior([st.norm(1, 2), st.norm(-1, 0.5), st.uniform(0, 1)], pot)
        tr = drivers.sample(post, draws=400, tune=300, chains=16, seed=3, sampler=smp, max_leapfrog=16)
        x = np.stack([tr.posterior[f'x{k}'] for k in range(3)], -1).reshape(-1, 3)
        assert np.allclose(x.mean(0), np.r_[mean, 0.5], atol=0.06), (smp, x.mean(0))
        assert np.allclose(x.var(0), np.r_[np.diag(cov), 1 / 12], rtol=0.15), (smp, x.var(0))
        assert abs(np.cov(x[:, 0], x[:, 1])[0, 1] - cov[0, 1]) < 0.04
        assert 0.6 < tr.sample_stats['acceptance_rate'].mean() < 0.95
        if smp == 'nuts':
            n = tr.sample_stats['n_steps']
            assert n.min() >= 1 and n.max() <= 1023 and np.all(tr.sample_stats['tree_depth'] <= 10)
            assert len(np.unique(n)) > 1            # chains stop at different depths within one lock-step transition


def test_gpmcmc_pickles_without_device_state(tmp_path):
    """save_object / load_object (andvaranaut/core.py) on a surrogate: the engine handle, the cached factorisation and
    the process-group shard are dropped; data, conversions and fitted hypers survive."""
    from andvaranaut_b200 import GPMCMC, save_object, load_object, meanstd
    g = GPMCMC(kernel='Matern52', noise=True, nx=2, ny=1, priors=[st.uniform(0, 1)] * 2, target=lambda x: np.array([x[0] + x[1]]),
               verbose=False, rundir=str(tmp_path / 'runs'))
    rng = np.random.default_rng(0)
    x = rng.uniform(size=(12, 2))
    g.set_data(x, (x[:, 0] + x[:, 1])[:, None])
    g.change_yconrevs([meanstd(g.y[:, 0])])
    g.hypers = {'gv': np.array(1e-4), 'l': np.array([1.0, 2.0]), 'kv': np.array([1.5])}
    g.gp = object()                 # stands for a GPEngine (ctypes handle + device tensors: not picklable state)
    g._pred_cache = dict(jitter=1e-6, theta=None, n=12, sum=None, eng=g.gp)
    f = str(tmp_path / 'g.pickle')
    save_object(g, f)
    h = load_object(f)
    assert h.gp is None and h._pred_cache is None and h.shard is None
    assert np.array_equal(h.x, g.x) and np.array_equal(h.yc, g.yc) and np.array_equal(h.hypers['l'], g.hypers['l'])
    assert g.gp is not None       # the live object keeps its device state


def test_batched_restarts_propagate_an_evaluation_failure_without_hanging():
    """a device / collective error inside the batched evaluation reaches every optimiser thread and ends the fit with
    that error (the rendezvous used to leave the other threads waiting forever)."""
    class Boom(RuntimeError):
        pass

    class FailingPosterior:
        def __init__(self):
            self.n = 0

        def logp_dlogp(self, z, jacobian):
            self.n += 1
            if self.n == 3:
                raise Boom('device lost')
            z = np.atleast_2d(z)
            return -0.5 * np.sum(z * z, axis=1), -z, np.zeros(len(z), dtype=np.int32)

    import threading
    res = {}

    def run():
        try:
            drivers.find_map_multi(FailingPosterior(), np.random.default_rng(0).normal(size=(4, 3)))
            res['out'] = 'returned'
        except Boom as e:
            res['out'] = e
    t = threading.Thread(target=run, daemon=True)
    t.start()
    t.join(timeout=60)
    assert not t.is_alive(), 'find_map_multi hung after a failed evaluation'
    assert isinstance(res['out'], Boom)


def test_theta_only_equals_theta_from_z():
    """the drivers launch the device evaluation on ParamSpace.theta_only(z) and form the Jacobian terms afterwards: it
    must be the first output of theta_from_z bit for bit, for every prior family / transform and for batches."""
    from andvaranaut_b200.priors import ParamSpace
    rng = np.random.default_rng(5)
    for truncate in (False, True):
        for kw in (dict(nx=3, nkern=1, noise=True),
                   dict(nx=8, nkern=1, noise=True, n_iw=16, cw_pos=[False, True, False, True]),
                   dict(nx=2, nkern=3, noise=False, has_alpha=True)):
            sp = ParamSpace(truncate=truncate, **kw)
            z = rng.normal(size=(7, sp.P)) * 2.0
            assert np.array_equal(sp.theta_only(z), sp.theta_from_z(z)[0])
            assert np.array_equal(sp.theta_only(z[0]), sp.theta_from_z(z[0])[0])

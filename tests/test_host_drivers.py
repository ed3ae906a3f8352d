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

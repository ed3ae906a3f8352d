"""GPU tests of the drop-in surface: the GPMCMC class end to end (tutorial workflow, warped fits, BO, MCMC) with
the oracle as the checker at the fitted hyperparameters."""
import os
import sys

import numpy as np
import pytest
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import gp_oracle as go  # noqa: E402
from andvaranaut_b200 import GPMCMC, maxmin, meanstd, uniform, normal, wgp  # noqa: E402
from andvaranaut_b200 import drivers  # noqa: E402
from andvaranaut_b200.priors import ParamSpace  # noqa: E402
from fake_engine import OracleEngine  # noqa: E402


def target_fun(x):
    # tutorial target function (tutorial.ipynb:61-63)
    x1, x2 = x
    return np.array([x1 ** 2 - x1 - x2 ** 2 * x1 + x2])


SPACE = [st.uniform(loc=0, scale=2), st.uniform(loc=1, scale=0.5)]


def tutorial_gp(tmp_path, kernel='RBF', noise=False, n=100):
    g = GPMCMC(kernel=kernel, noise=noise, xconrevs=[uniform(SPACE[0]), normal(SPACE[1])], yconrevs=[None],
               nx=2, ny=1, priors=SPACE, target=target_fun, parallel=False, nproc=1, verbose=False,
               rundir=str(tmp_path / 'runs'))
    g.sample(n, seed=101)
    g.change_conrevs([maxmin(g.x[:, 0]), maxmin(g.x[:, 1])], [meanstd(g.y[:, 0])])
    return g


def test_tutorial_workflow_config1(tmp_path):
    """C1: LHC N=100, d=2, RBF, noise=False, MAP fit, predict (tutorial.ipynb cells 18-30)."""
    g = tutorial_gp(tmp_path)
    assert g.x.shape == (100, 2) and g.xc.min() >= 0.0099 and g.xc.max() <= 0.9901
    g.fit(restarts=1)
    h = g.hypers
    assert set(h) == {'l_log__', 'kv_log__', 'l', 'kv'}           # tutorial.ipynb:529
    assert np.allclose(np.exp(h['l_log__']), h['l']) and h['l'].shape == (2,) and h['kv'].shape == (1,)
    # plausibility band of the tutorial fit: l ~ O(1), kv ~ O(10-100)
    assert 0.3 < h['l'].min() and h['l'].max() < 20 and 1.0 < h['kv'][0] < 1e4
    # the MAP is a stationary point of the ORACLE posterior too
    spec = go.ModelSpec(nx=2, kerns=['RBF'], noise=False)
    eng = OracleEngine(spec)
    eng.set_data(g.xc, g.yc[:, 0])
    sp = ParamSpace(2, 1, False)
    post = drivers.Posterior(eng, sp)
    z = np.concatenate([h['l_log__'], h['kv_log__']])
    v, grad, _ = post.logp_dlogp(z[None, :], False)
    assert np.max(np.abs(grad)) < 5e-2 and 380 < v[0] < 560        # tutorial logp: 466.88
    # predictions: device vs oracle at the fitted hypers, and accuracy band of the tutorial (RMSE ~1e-4)
    rng = np.random.default_rng(0)
    xt = np.c_[rng.uniform(0.05, 1.95, 200), rng.uniform(1.02, 1.48, 200)]
    y, yv = g.predict(xt, return_var=True)
    truth = np.array([target_fun(x)[0] for x in xt])
    assert y.shape == (200, 1) and np.sqrt(np.mean((y[:, 0] - truth) ** 2)) < 2e-3
    th = sp.theta_from_hypers(h)
    xct = np.c_[g.xconrevs[0].con(xt[:, 0]), g.xconrevs[1].con(xt[:, 1])]
    mu_r, var_r = go.predict(spec, th, g.xc, g.yc[:, 0], xct)
    m_ref, v_ref = go.gh_stats_loop(mu_r, var_r, g.yconrevs[0].rev, normvar=False)
    assert np.max(np.abs(y - m_ref)) <= 1e-6 * np.max(np.abs(m_ref))      # cond(K) ~ 1e9 at these hypers
    # held-out metrics as test_plots prints them
    g.train_test(0.9, seed=1)
    met = g.test_metrics()
    assert met['rmse'] < 5e-3 and met['r2'] > 0.9999


def test_matern_noise_fit_and_restarts(tmp_path):
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=60)
    g.fit(method='map', restarts=3, seed=4)
    assert set(g.hypers) == {'gv_log__', 'l_log__', 'kv_log__', 'gv', 'l', 'kv'}
    assert g.hypers['gv'].shape == () and 0 < float(g.hypers['gv']) < 1e-2
    y = g.predict(g.x[:10])
    assert np.max(np.abs(y[:, 0] - g.y[:10, 0])) < 5e-3


def test_warped_fit_bakes_parameters(tmp_path):
    """iwgp + cwgp: the learnt warp parameters are baked into NumPy conrevs and the converted caches refreshed
    (gpmcmc.py:364-399); the objective the device optimised equals the oracle's at the optimum."""
    rng = np.random.default_rng(3)
    pri = [st.uniform(0, 1)] * 3

    def f(x):
        return np.array([np.exp(np.sin(3 * x[0]) + x[1] * x[2])])
    xcon = [wgp(['uniform', 'kumaraswamy'], [1.0, 1.0], xdist=pri[i]) for i in range(3)]
    g = GPMCMC(kernel='Matern52', noise=True, xconrevs=xcon, yconrevs=[None], nx=3, ny=1, priors=pri, target=f,
               verbose=False, rundir=str(tmp_path / 'runs'))
    g.sample(80, seed=7)
    g.change_yconrevs([wgp(['logarithm', 'sal', 'meanstd'], [0.0, 1.0, 0.0, 1.0], y=g.y[:, 0])])
    data = g.fit(method='map', iwgp=True, cwgp=True, return_data=True)
    h = g.hypers
    assert h['iwgp'].shape == (6,) and h['cwgp_pos'].shape == (2,) and h['cwgp'].shape == (2,)
    assert np.allclose(g.xconrevs[0].params, h['iwgp'][:2])
    assert np.allclose(g.yconrevs[0].params, [h['cwgp'][0], h['cwgp_pos'][0], h['cwgp'][1], h['cwgp_pos'][1]])
    assert np.allclose(g.xc[:, 1], g.xconrevs[1].con(g.x[:, 1])) and np.allclose(g.yc[:, 0], g.yconrevs[0].con(g.y[:, 0]))
    spec = go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 3,
                        ywarp=['logarithm', 'sal', 'meanstd'])
    eng = OracleEngine(spec)
    eng.set_data(g.x, g.y[:, 0])
    sp = ParamSpace(3, 1, True, n_iw=6, cw_pos=[False, True, False, True])
    z = sp.z_from_theta(sp.theta_from_hypers(h))
    v, grad, _ = drivers.Posterior(eng, sp).logp_dlogp(z[None, :], False)
    assert abs(v[0] - data['logp']) <= 1e-8 * abs(v[0])
    yp = g.predict(g.x[:20])
    assert np.max(np.abs(yp[:, 0] - g.y[:20, 0]) / np.abs(g.y[:20, 0])) < 0.05


def test_bayesian_optimisation_improves_optimum(tmp_path):
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=False, n=24)
    g.fit()
    y0 = g.y.min()
    n0 = len(g.x)
    xopt, yopt = g.BO(opt_type='min', max_iter=4, predict_samps=2048, seed=3)
    assert len(g.x) > n0 and len(g.xc) == len(g.x) and len(g.ym) == len(g.x)
    assert yopt <= y0 and np.isclose(yopt, g.y.min())
    true_min = min(target_fun(np.array([a, b]))[0] for a in np.linspace(0, 2, 201) for b in np.linspace(1, 1.5, 51))
    assert yopt < true_min + 0.05


def test_mcmc_fit_methods(tmp_path):
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=40)
    data = g.fit(method='mcmc_mean', draws=60, tune=100, chains=8, seed=2, max_leapfrog=10, return_data=True)
    assert data.posterior['l'].shape == (8, 60, 2) and np.isfinite(data.sample_stats['lp']).all()
    assert np.allclose(g.hypers['l'], data.posterior['l'].mean(axis=(0, 1)))
    y = g.predict(g.x[:5])
    assert np.max(np.abs(y[:, 0] - g.y[:5, 0])) < 2e-2
    g.fit(method='mcmc_map', draws=30, tune=60, chains=4, seed=2, max_leapfrog=8)
    assert set(g.hypers) >= {'gv', 'l', 'kv'}


def test_predict_with_mean_function_and_user_transform(tmp_path):
    """non-zero mean (evaluated on the host, added inside the GH epilogue) and a transform the device cannot
    evaluate (host reversion)."""
    class user_asinh:
        def con(self, y):
            return np.arcsinh(2.0 * y)

        def rev(self, y):
            return 0.5 * np.sinh(y)
    g = GPMCMC(kernel='RBF', noise=True, xconrevs=[uniform(SPACE[0]), uniform(SPACE[1])], yconrevs=[user_asinh()],
               mean=lambda x: np.array([x[0] - 1.0]), nx=2, ny=1, priors=SPACE, target=target_fun, verbose=False,
               rundir=str(tmp_path / 'runs'))
    g.sample(50, seed=11)
    g.fit()
    y = g.predict(g.x[:8])
    assert np.max(np.abs(y[:, 0] - g.y[:8, 0])) < 0.05


@pytest.mark.parametrize('method,opt_type', [('EI', 'min'), ('EI', 'max'), ('exploit', 'min'), ('explore', 'min')])
def test_acquisition_grad_matches_oracle_graph(tmp_path, method, opt_type):
    """the BO refine graph (gpmcmc.py:738-829): value against the oracle's restatement of it (variance WITHOUT gv,
    GH reversion, EI), gradient w.r.t. the RAW query point against central differences of that oracle value."""
    g = GPMCMC(kernel='Matern52', noise=True, xconrevs=[uniform(SPACE[0]), normal(SPACE[1])],
               yconrevs=[None], mean=lambda x: np.array([0.1 * x[0]]),
               nx=2, ny=1, priors=SPACE, target=lambda x: target_fun(x) + 3.0, verbose=False, rundir=str(tmp_path / 'runs'))
    g.sample(40, seed=7)
    g.change_yconrevs([wgp(['logarithm', 'sal', 'meanstd'], [0.0, 1.0, 0.0, 1.0], y=(g.y - g.ym)[:, 0])])
    g.fit(cwgp=True)
    g.yopt = float(np.min(g.y)) if opt_type == 'min' else float(np.max(g.y))
    spec = go.ModelSpec(nx=2, kerns=['Matern52'], noise=True)
    sp = ParamSpace(2, 1, True)
    th = sp.theta_from_hypers(g.hypers)
    w = g.yconrevs[0]

    def oracle_value(x):
        xc = np.column_stack([g.xconrevs[i].con(x[:, i]) for i in range(2)])
        mu, var = go.predict(spec, th, g.xc, g.yc[:, 0], xc)
        var = var - go.unpack(spec, th)['gv']
        madd = 0.1 * x[:, 0]
        m, v = go.gh_stats(mu, var, w.rev, mean_add=madd, normvar=(method == 'explore'), EI=(method == 'EI'),
                           EIopt=opt_type, yopt=g.yopt)
        if method == 'explore':
            return -v[:, 0]
        if method == 'EI':
            return -m[:, 0]
        return m[:, 0] if opt_type == 'min' else -m[:, 0]

    rng = np.random.default_rng(5)
    x = np.column_stack([rng.uniform(0.1, 1.9, 12), rng.uniform(1.05, 1.45, 12)])
    f, gr = g.acquisition_grad(x, method=method, opt_type=opt_type, normvar=True)
    ref = oracle_value(x)
    assert np.max(np.abs(f - ref)) <= 1e-8 * np.max(np.abs(ref))
    # the normalised variance is ~1e-6 and a cancelling difference (kv - |v|^2): its central differences carry
    # ~1e-10 / h of rounding noise, hence the wider step and tolerance for 'explore'
    hh, tol = (1e-5, 1e-3) if method == 'explore' else (1e-6, 2e-5)
    for i in range(2):
        h = np.zeros(2)
        h[i] = hh
        fd = (oracle_value(x + h) - oracle_value(x - h)) / (2 * hh)
        assert np.max(np.abs(gr[:, i] - fd)) <= tol * max(np.max(np.abs(fd)), np.max(np.abs(gr))), (i, gr[:, i], fd)


def test_nuts_fit_on_device(tmp_path):
    """fit(method='mcmc_mean') with the lock-step NUTS driver (default sampler): every leapfrog of every chain is one
    batched device call; trajectory lengths differ between chains, draws are finite and predictive."""
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=40)
    data = g.fit(method='mcmc_mean', draws=40, tune=80, chains=8, seed=3, sampler='nuts', max_treedepth=6,
                 return_data=True)
    n = data.sample_stats['n_steps']
    assert n.shape == (8, 40) and n.min() >= 1 and n.max() <= 63 and len(np.unique(n)) > 1
    assert np.isfinite(data.sample_stats['lp']).all() and data.sample_stats['acceptance_rate'].mean() > 0.5
    y = g.predict(g.x[:5])
    assert np.max(np.abs(y[:, 0] - g.y[:5, 0])) < 2e-2


@pytest.mark.parametrize('with_var', [False, True])
def test_inverse_posterior_matches_oracle_graph(tmp_path, with_var):
    """inverse_opt's model (gpmcmc.py:1049-1165) on the device (one factorisation + batched predict_grad, block form)
    against the oracle's literal dense restatement: logp over the unconstrained x to 1e-9, gradient to central
    differences of the oracle."""
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=60)
    g.fit()
    yobs = np.array([[0.35]])
    yvar = np.array([[1e-3]]) if with_var else None
    post = g.inverse_posterior(yobs, yvar)
    spec = go.ModelSpec(nx=2, kerns=['Matern52'], noise=True, jitter=0.0)
    sp = ParamSpace(2, 1, True)
    th = sp.theta_from_hypers(g.hypers)
    c = g.yconrevs[0]
    noise_o = np.sqrt(go.gh_stats_inv(yobs, yvar, c.con)) if with_var else 0.0
    ynoise = np.r_[np.full(60, np.sqrt(g.hypers['gv'] + 1e-6)), noise_o]
    lyd = np.sum(np.log(c.der(np.r_[g.y[:, 0], yobs[:, 0]])))

    def oracle_logp(z):
        x = post.space.theta_from_z(z)[0]
        xc = np.array([g.xconrevs[i].con(x[i:i + 1])[0] for i in range(2)])
        return go.inverse_loglik(spec, th, g.xc, g.yc[:, 0], xc, c.con(yobs[:, 0]), ynoise, lyd) \
            + post.space.prior(x)[0]

    rng = np.random.default_rng(6)
    z = rng.normal(size=(9, 2))
    v, gr, _ = post.logp_dlogp(z, False)
    ref = np.array([oracle_logp(zz) for zz in z])
    assert np.max(np.abs(v - ref)) <= 1e-9 * np.max(np.abs(ref)), (v, ref)
    # e^2 / s with a small predictive variance s: the dense oracle's central differences carry ~1e-10 / h of rounding
    # noise relative to gradients of O(1e3), so the step is 1e-4 (truncation ~1e-7, measured on the CPU twin of this test)
    for i in range(2):
        h = np.zeros(2)
        h[i] = 1e-4
        fd = np.array([(oracle_logp(zz + h) - oracle_logp(zz - h)) / 2e-4 for zz in z])
        assert np.max(np.abs(gr[:, i] - fd)) <= 5e-5 * max(1.0, np.max(np.abs(fd))), (gr[:, i], fd)


def test_inverse_opt_recovers_an_input_that_explains_the_observation(tmp_path):
    g = tutorial_gp(tmp_path, kernel='RBF', noise=True, n=80)
    g.fit()
    xtrue = np.array([1.3, 1.2])
    yobs = np.array([target_fun(xtrue)])
    data, xopt = g.inverse_opt(yobs, method='map', restarts=6, seed=1)
    assert xopt.shape == (2,) and set(data) == {'x0', 'x1', 'x0_interval__', 'x1_interval__'}
    assert abs(g.predict(xopt[None, :])[0, 0] - yobs[0, 0]) < 5e-3
    n0 = len(g.x)
    data, xopt, ysamp = g.inverse_opt(yobs, yvarobs=np.array([[1e-4]]), method='mcmc_mean', evaluate_opt=True,
                                      draws=40, tune=60, chains=4, seed=2)
    assert data.posterior['x0'].shape == (4, 40) and len(g.x) == n0 + 1 and len(g.xc) == n0 + 1
    assert np.allclose(g.x[-1], xopt) and np.isfinite(ysamp).all()


def test_bo_with_the_acquisition_model(tmp_path):
    """BO(opt_method='map'): candidates from find_MAP over the x model with the device acquisition as Potential
    (gpmcmc.py:699-858), instead of LHC candidates."""
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=30)
    g.fit()
    y0 = float(np.min(g.y))
    xopt, yopt = g.BO(opt_type='min', opt_method='map', max_iter=3, method='EI', seed=4, restarts=4)
    assert len(g.x) > 30 and yopt <= y0


def test_appended_points_extend_the_cached_factorisation(tmp_path):
    """BO with fit_method='none' (hypers reused, gpmcmc.py:347-349): the new samples extend the cached factorisation by
    rank-1 appends; predictions equal those of a surrogate built from scratch on the enlarged data set."""
    g = tutorial_gp(tmp_path, kernel='Matern52', noise=True, n=50)
    g.fit()
    g.predict(g.x[:3])                       # factorise + cache
    eng0 = g._pred_cache['eng']
    g.BO(opt_type='min', max_iter=3, method='EI', fit_method='none', predict_samps=500, seed=1, refine=False)
    assert len(g.x) == 53
    xq = np.column_stack([np.linspace(0.1, 1.9, 25), np.linspace(1.05, 1.45, 25)])
    m1, v1 = g.predict(xq, return_var=True)
    assert g._pred_cache['eng'] is eng0 and eng0.N == 53          # same engine, extended in place
    g._pred_cache = None
    m2, v2 = g.predict(xq, return_var=True)                   # refactorised from scratch
    assert g._pred_cache['eng'] is not eng0
    assert np.max(np.abs(m1 - m2)) <= 1e-8 * np.max(np.abs(m2))
    assert np.max(np.abs(v1 - v2)) <= 1e-8 * np.max(np.abs(v2)) + 1e-12


def test_y_dist_returns_surrogate_samples(tmp_path):
    g = tutorial_gp(tmp_path, kernel='RBF', noise=True, n=40)
    g.fit()
    xs, ys = g.y_dist(nsamps=300, return_data=True, seed=3)
    assert xs.shape == (300, 2) and ys.shape == (300, 1)
    truth = np.array([target_fun(x)[0] for x in xs])
    assert np.sqrt(np.mean((ys[:, 0] - truth) ** 2)) < 0.05
    with pytest.raises(Exception, match='mode must be one of'):
        g.y_dist(mode='violin', nsamps=10)

"""CPU tests: the oracle against the reference-generated fixtures, its own finite differences and
scipy.stats; the product's host transforms against the same fixtures."""
import os
import sys

import numpy as np
import pytest
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))

from oracle import gp_oracle as go  # noqa: E402
from oracle.warp_oracle import WarpOracle, Dual  # noqa: E402
from andvaranaut_b200 import transform as T  # noqa: E402
import make_golden as mg  # noqa: E402
import cases  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'transforms_ref.npz'))


def close(a, b, tol=1e-13):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)) <= tol


def test_tutorial_known_answers():
    # tutorial.ipynb:362-369 -- uniform/normal conversion of the first two LHC samples and the target fn
    k = np.load(os.path.join(HERE, 'golden', 'tutorial_kat.npz'))
    space = [st.uniform(loc=0, scale=2), st.uniform(loc=1, scale=0.5)]
    got = np.c_[T.uniform(space[0]).con(k['x'][:, 0]), T.normal(space[1]).con(k['x'][:, 1])]
    assert np.allclose(got, k['xc'], atol=5e-9)
    wo = WarpOracle(['uniform'], [], xdist_interval=(0.0, 2.0))
    assert np.allclose(wo.con(k['x'][:, 0]), k['xc'][:, 0], atol=5e-9)


def test_fixed_transforms_match_reference():
    assert close(T.uniform(st.uniform(0, 2)).con(G['x_u']), G['uniform_con'])
    assert close(T.normal(st.uniform(1, 0.5)).con(G['x_n']), G['normal_con'])
    assert close(T.maxmin(G['x_u']).con(G['x_u']), G['maxmin_con'])
    mm = T.maxmin(G['x_u'])
    assert close([mm.a, mm.b], G['maxmin_ab'])
    assert close(T.maxmin(G['x_u'], centred=True).con(G['x_u']), G['maxminc_con'])
    ms = T.meanstd(G['y'])
    assert close(ms.con(G['y']), G['meanstd_con']) and close(ms.rev(ms.con(G['y'])), G['meanstd_rev'])
    for nm, cls in [('probit', T.probit), ('cdf', T.cdf), ('logit_logistic', T.logit_logistic)]:
        c = cls(st.norm(1.0, 2.0))
        assert close(c.con(G['y']), G[nm + '_con'], 1e-12)
        assert close(c.rev(c.con(G['y'])), G[nm + '_rev'], 1e-12)
    for nm, cls in [('nonneg', T.nonneg), ('log1p', T.log1p), ('log10', T.log10)]:
        c = cls()
        assert close(c.con(G['ypos']), G[nm + '_con'], 1e-12)
        assert close(c.rev(c.con(G['ypos'])), G[nm + '_rev'], 1e-12)
    assert close(T.normalise(3.5).con(G['y']), G['normalise_con'])


@pytest.mark.parametrize('case', mg.WGP_CASES, ids=[c[0] for c in mg.WGP_CASES])
def test_wgp_matches_reference(case):
    name, stages, params, kind, interval = case
    d = G[f'wgp_{name}_data']
    xd = st.uniform(interval[0], interval[1] - interval[0]) if interval else None
    w = T.wgp(stages, params, y=d, xdist=xd)
    assert close(w.con(d), G[f'wgp_{name}_con'])
    assert close(w.der(d), G[f'wgp_{name}_der'], 1e-12)
    assert close(w.rev(w.con(d)), G[f'wgp_{name}_rev'])
    t = G[f'wgp_{name}_test']
    assert close(w.con(t), G[f'wgp_{name}_test_con']) and close(w.der(t), G[f'wgp_{name}_test_der'], 1e-12)
    assert (w.pos.astype(np.int8) == G[f'wgp_{name}_pos']).all() and (w.pid == G[f'wgp_{name}_pid']).all()
    assert w.np == int(G[f'wgp_{name}_np'])
    # oracle, both modes
    for duals in (False, True):
        wo = WarpOracle(stages, params, y=d, xdist_interval=interval, with_duals=duals)
        z = wo._ycon.v if duals else wo._ycon
        assert close(z, G[f'wgp_{name}_con'])
        dr = wo.der(Dual.lift(d, len(params))) if duals else wo.der(d)
        assert close(dr.v if duals else dr, G[f'wgp_{name}_der'], 1e-12)
    assert (wo.pos == w.pos).all() and (wo.pid == w.pid).all() and wo.np == w.np


def test_wgp_duals_match_finite_differences():
    rng = np.random.default_rng(5)
    for name, stages, params, kind, interval in mg.WGP_CASES:
        d = mg.data_of(kind, rng, 30)
        p = np.array(params, dtype=np.float64)
        wo = WarpOracle(stages, p, y=d, xdist_interval=interval, with_duals=True)
        for q in range(len(p)):
            h = 1e-6 * max(1.0, abs(p[q]))
            pp, pm = p.copy(), p.copy()
            pp[q] += h
            pm[q] -= h
            zp = WarpOracle(stages, pp, y=d, xdist_interval=interval)._ycon
            zm = WarpOracle(stages, pm, y=d, xdist_interval=interval)._ycon
            fd = (zp - zm) / (2 * h)
            assert np.max(np.abs(fd - wo._ycon.d[:, q])) <= 1e-6 * max(1.0, np.max(np.abs(fd))), (name, q)


def fd5(spec, theta, X, y, h=2e-4):
    g = np.zeros_like(theta)

    def f(t):
        return go.loglik(spec, t, X, y, False).ll
    for i in range(len(theta)):
        hh = h * abs(theta[i]) if abs(theta[i]) > 1e-8 else h
        e = np.zeros_like(theta)
        e[i] = hh
        g[i] = (-f(theta + 2 * e) + 8 * f(theta + e) - 8 * f(theta - e) + f(theta - 2 * e)) / (12 * hh)
    return g


SPECS = {
    'rbf': go.ModelSpec(nx=3, kerns=['RBF']),
    'm52': go.ModelSpec(nx=3, kerns=['Matern52']),
    'm32': go.ModelSpec(nx=3, kerns=['Matern32']),
    'expo': go.ModelSpec(nx=3, kerns=['Exponential']),
    'rq': go.ModelSpec(nx=3, kerns=['RatQuad']),
    'sum': go.ModelSpec(nx=3, kerns=['RBF', 'Matern52'], ops=['+']),
    'mix3': go.ModelSpec(nx=2, kerns=['RBF', 'Matern32', 'Exponential'], ops=['*', '+']),
    'rqprod': go.ModelSpec(nx=2, kerns=['RatQuad', 'Matern52'], ops=['*']),
    'warps': go.ModelSpec(nx=3, kerns=['Matern52'], xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0)), None,
                                                              (['kumaraswamy', 'maxmin'], None)],
                          ywarp=['logarithm', 'sal', 'meanstd']),
    'ywarp_all': go.ModelSpec(nx=2, kerns=['RBF'], ywarp=['affine', 'arcsinh', 'boxcox', 'sinharcsinh', 'stdshift',
                                                            'pzero']),
}


@pytest.mark.parametrize('name', list(SPECS))
def test_oracle_gradient_vs_finite_differences(name):
    spec = SPECS[name]
    X, y, th, _ = cases.synth(spec, 40, seed=11)
    r = go.loglik(spec, th, X, y)
    g = fd5(spec, th, X, y)
    scale = np.abs(g) + 1e-4 * np.max(np.abs(g))
    # Exponential: dk/dr2 = -k/(4r) is ~2.5e5 on the diagonal (r = 1e-6), so rounding noise of the gram-form
    # r2_ii (a few 1e-16) makes the finite differences themselves noisy at the 1e-5 level
    tol = 1e-4 if 'Exponential' in spec.kerns else 2e-6
    assert np.max(np.abs(r.grad - g) / scale) < tol, (r.grad, g)


def test_oracle_golden_fixtures_do_not_drift():
    for name, case in mg.gp_cases().items():
        g = np.load(os.path.join(HERE, 'golden', f'gp_oracle_{name}.npz'))
        X, y, th, Xs = mg.gp_inputs(case)
        assert np.array_equal(X, g['X']) and np.array_equal(th, g['theta'])
        r = go.loglik(case['spec'], th, X, y, keep=True)
        assert abs(r.ll - float(g['ll'])) <= 1e-10 * abs(float(g['ll']))
        assert np.allclose(r.grad, g['grad'], rtol=1e-7, atol=1e-7 * np.max(np.abs(g['grad'])))
        if case['M']:
            mu, var = go.predict(case['spec'], th, r.Xw, r.z, Xs)
            assert np.allclose(mu, g['mu'], rtol=1e-8, atol=1e-9) and np.allclose(var, g['var'], rtol=1e-7, atol=1e-9)
            nr = mg.next_row_outputs(case['spec'], th, r.Xw, r.z, Xs)
            for k, v in nr.items():
                assert np.allclose(v, g[k], rtol=1e-7, atol=1e-9 * max(1.0, float(np.max(np.abs(g[k]))))), (name, k)


def test_tutorial_plausibility_band():
    # tutorial.ipynb:488,529 -- RBF, N=100, d=2, noise=False: logp ~ +467 at l~(1.13,2.69), kv~68
    g = np.load(os.path.join(HERE, 'golden', 'gp_oracle_rbf_c1.npz'))
    lp = (float(g['ll']) + np.sum(go.logp_lognormal(g['theta'][:2], 0.0, 1.0))
          + go.logp_lognormal(g['theta'][2], 0.56, 0.75))
    assert 400 < lp < 520


def test_predict_blocked_equals_predict():
    spec = go.ModelSpec(nx=3, kerns=['Matern52'])
    X, y, th, Xs = cases.synth(spec, 50, seed=3, M=37)
    mu, var = go.predict(spec, th, X, y, Xs)
    mu2, var2, _ = go.predict_blocked(spec, th, X, y, Xs, block=16)
    assert np.allclose(mu, mu2, rtol=1e-12) and np.allclose(var, var2, rtol=1e-10, atol=1e-14)


def test_gh_stats_vectorised_equals_loop():
    rng = np.random.default_rng(2)
    mu, var = rng.normal(size=25), rng.uniform(0.01, 0.5, 25)
    w = T.wgp(['logarithm', 'sal', 'meanstd'], [0.1, 1.1, -0.2, 0.9], y=np.exp(rng.normal(size=40)))
    madd = rng.normal(size=25)
    for kw in [dict(normvar=False), dict(normvar=True), dict(EI=True, EIopt='max', yopt=1.0, normvar=False),
               dict(EI=True, EIopt='min', yopt=1.5, normvar=False)]:
        a = go.gh_stats_loop(mu, var, w.rev, mean_add=madd, **kw)
        b = go.gh_stats(mu, var, w.rev, mean_add=madd, **kw)
        assert np.allclose(a[0], b[0], rtol=1e-13) and np.allclose(a[1], b[1], rtol=1e-11, atol=1e-13)


def test_prior_logps_match_scipy():
    x = np.array([0.3, 1.0, 2.5])
    assert np.allclose(go.logp_lognormal(x, 0.56, 0.75), st.lognorm(s=0.75, scale=np.exp(0.56)).logpdf(x))
    assert np.allclose(go.logp_halfnormal(x * 1e-3, 1e-3), st.halfnorm(scale=1e-3).logpdf(x * 1e-3))
    assert np.allclose(go.logp_normal(x, 0.0, 1.0), st.norm().logpdf(x))
    a, b = (1e-3 - 0.5) / 0.15, (100 - 0.5) / 0.15
    assert np.allclose(go.logp_truncnormal(x, 0.5, 0.15, 1e-3, 100.0), st.truncnorm(a, b, loc=0.5, scale=0.15).logpdf(x))


# ---- independent pin of the GP algebra: scikit-learn's GaussianProcessRegressor (installed here) ----------------
# PyMC itself cannot be installed offline, so the oracle's kernel / marginal-likelihood / conditional formulas are
# checked against a second, unrelated implementation of the same textbook algebra.  sklearn parameterises in log
# space (gradient = theta * d/dtheta) and evaluates Matern kernels at the exact distance, PyMC at sqrt(r2 + 1e-12):
# the two differ by ~1e-12 relative on the diagonal, hence the 1e-8 tolerances for the Matern cases.
@pytest.mark.parametrize('kern,tol', [('RBF', 1e-10), ('Matern52', 1e-8), ('Matern32', 1e-8)])
def test_oracle_matches_scikit_learn_gp(kern, tol):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, Matern, ConstantKernel, WhiteKernel
    d, N, M = 3, 60, 25
    spec = go.ModelSpec(nx=d, kerns=[kern], noise=True, jitter=1e-6)
    X, y, th, Xs = cases.synth(spec, N, seed=31, M=M)
    p = go.unpack(spec, th)
    base = RBF(length_scale=p['l']) if kern == 'RBF' else Matern(length_scale=p['l'], nu=2.5 if kern == 'Matern52' else 1.5)
    k = ConstantKernel(p['kv'][0]) * base + WhiteKernel(p['gv'])
    gpr = GaussianProcessRegressor(kernel=k, alpha=spec.jitter, optimizer=None).fit(X, y)
    lml, glog = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    r = go.loglik(spec, th, X, y)
    assert abs(r.ll - lml) <= tol * abs(lml)
    # sklearn's theta order: log kv, log l[0..d-1], log gv
    o = spec.offsets()
    g_sk = np.concatenate([[glog[-1] / p['gv']], glog[1:1 + d] / p['l'], [glog[0] / p['kv'][0]]])
    g_or = np.concatenate([[r.grad[o['gv']]], r.grad[o['l']:o['l'] + d], [r.grad[o['kv']]]])
    assert np.max(np.abs(g_or - g_sk) / np.maximum(np.abs(g_sk), 1e-3 * np.max(np.abs(g_sk)))) <= max(tol, 1e-9) * 10
    mu_sk, sd_sk = gpr.predict(Xs, return_std=True)           # includes the WhiteKernel term: pred_noise=True
    mu, var = go.predict(spec, th, X, y, Xs)
    assert np.max(np.abs(mu - mu_sk)) <= 1e-7 * np.max(np.abs(mu_sk))
    assert np.max(np.abs(var - sd_sk ** 2) / np.maximum(sd_sk ** 2, p['kv'][0])) <= 1e-7


# ---- BO refine graph: gradients of the predictive mean / variance w.r.t. the query points -------------------
@pytest.mark.parametrize('name', ['rbf', 'm52', 'm32', 'rq', 'sum', 'mix3', 'rqprod'])
def test_oracle_predict_grad_vs_finite_differences(name):
    spec = SPECS[name]
    X, y, th, Xs = cases.synth(spec, 40, seed=21, M=6)
    for pn in (True, False):
        mu, var, dmu, dvar = go.predict_grad(spec, th, X, y, Xs, pred_noise=pn)
        mu0, var0 = go.predict(spec, th, X, y, Xs)
        gv = go.unpack(spec, th)['gv']
        assert np.allclose(mu, mu0, rtol=1e-12) and np.allclose(var, var0 - (0.0 if pn else gv), rtol=1e-10, atol=1e-13)
    h = 1e-6
    for m in range(spec.nx):
        e = np.zeros(spec.nx)
        e[m] = h
        mp, vp = go.predict(spec, th, X, y, Xs + e)
        mm, vm = go.predict(spec, th, X, y, Xs - e)
        assert np.allclose((mp - mm) / (2 * h), dmu[:, m], rtol=2e-6, atol=2e-6 * np.max(np.abs(dmu)))
        assert np.allclose((vp - vm) / (2 * h), dvar[:, m], rtol=2e-6, atol=2e-6 * np.max(np.abs(dvar)))


def test_oracle_gh_stats_grad_vs_finite_differences():
    rng = np.random.default_rng(4)
    M, d = 9, 3
    mu, var = rng.normal(size=M) * 0.5, rng.uniform(0.02, 0.4, M)
    dmu, dvar = rng.normal(size=(M, d)), 0.1 * rng.normal(size=(M, d))
    w = T.wgp(['logarithm', 'sal', 'meanstd'], [0.1, 1.1, -0.2, 0.9], y=np.exp(rng.normal(size=40)))
    madd, dmadd = rng.normal(size=M) * 0.1, rng.normal(size=(M, d)) * 0.1
    h = 1e-6
    for kw in [dict(normvar=False), dict(normvar=True), dict(EI=True, EIopt='max', yopt=0.7, normvar=False),
               dict(EI=True, EIopt='min', yopt=1.5, normvar=False)]:
        m, v, dm, dv = go.gh_stats_grad(mu, var, dmu, dvar, w.rev, w.der, mean_add=madd, dmean_add=dmadd, **kw)
        m0, v0 = go.gh_stats(mu, var, w.rev, mean_add=madd, **kw)
        assert np.allclose(m, m0[:, 0], rtol=1e-13) and np.allclose(v, v0[:, 0], rtol=1e-11, atol=1e-13)
        for k in range(d):   # move along direction k: latent mean / variance / mean function to first order
            a = go.gh_stats(mu + h * dmu[:, k], var + h * dvar[:, k], w.rev, mean_add=madd + h * dmadd[:, k], **kw)
            b = go.gh_stats(mu - h * dmu[:, k], var - h * dvar[:, k], w.rev, mean_add=madd - h * dmadd[:, k], **kw)
            assert np.allclose((a[0] - b[0])[:, 0] / (2 * h), dm[:, k], rtol=1e-6, atol=1e-7 * np.max(np.abs(dm))), kw
            assert np.allclose((a[1] - b[1])[:, 0] / (2 * h), dv[:, k], rtol=1e-6, atol=1e-7 * np.max(np.abs(dv))), kw


def test_oracle_against_extended_precision_truth():
    """oracle/gp_truth.py evaluates the same formulas in 50-digit arithmetic (fixtures: tests/golden/make_truth.py).  An
    independent pin of the oracle's algebra (kernels, likelihood, analytic gradient, conditional) and the measurement of
    its float64 noise floor: 1e-12 or better at ordinary conditioning; on the tutorial case (cond(K) ~ 6e9) the oracle
    itself is only good to ~5e-10 (ll) / ~1e-6 (gradient) -- the floor the C1 device test is held to."""
    g = np.load(os.path.join(HERE, 'golden', 'gp_truth_m52.npz'))
    spec = mg.gp_cases()['m52_noise']['spec']
    for b, t in enumerate(g['thetas']):
        r = go.loglik(spec, t, g['X'], g['y'])
        assert abs(r.ll - g['ll'][b]) <= 1e-12 * abs(g['ll'][b])
        assert np.max(np.abs(r.grad - g['grad'][b]) / np.maximum(np.abs(g['grad'][b]), 1e-3 * np.max(np.abs(g['grad'][b])))) <= 1e-11
    mu, var = go.predict(spec, g['thetas'][0], g['X'], g['y'], g['Xs'])
    assert np.max(np.abs(mu - g['mu'][0])) <= 1e-12 * np.max(np.abs(g['mu'][0]))
    assert np.max(np.abs(var - g['var'][0])) <= 1e-12 * g['thetas'][0][-1]
    c1 = np.load(os.path.join(HERE, 'golden', 'gp_truth_c1.npz'))
    spec1 = mg.gp_cases()['rbf_c1']['spec']
    e_ll = e_g = 0.0
    for b, t in enumerate(c1['thetas']):
        r = go.loglik(spec1, t, c1['X'], c1['y'])
        e_ll = max(e_ll, abs(r.ll - c1['ll'][b]) / abs(c1['ll'][b]))
        e_g = max(e_g, np.max(np.abs(r.grad - c1['grad'][b]) / np.maximum(np.abs(c1['grad'][b]), 1e-3 * np.max(np.abs(c1['grad'][b])))))
    assert 1e-12 < e_ll < 5e-9 and 1e-10 < e_g < 1e-5, (e_ll, e_g)     # the floor is real, and no worse than recorded


def test_truth_module_reproduces_its_fixture():
    """the committed truth fixture is what oracle/gp_truth.py produces (one hyperparameter vector, a few seconds)."""
    from oracle import gp_truth
    g = np.load(os.path.join(HERE, 'golden', 'gp_truth_m52.npz'))
    r = gp_truth.evaluate('Matern52', True, 1e-6, g['thetas'][1], g['X'], g['y'])
    assert r['ll'] == g['ll'][1] and np.array_equal(r['grad'], g['grad'][1])


def test_bench_golden_fixtures_are_the_oracle_on_the_bench_inputs():
    """drift guard for tests/golden/bench_*.npz: the benchmark's workload generators reproduce the inputs the fixtures
    were made from, and the oracle reproduces a row of each (c3 in full, c2 without the gradient: seconds)."""
    import zlib
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    for name, B, seed in (('c2', 64, 202), ('c3', 512, 303)):
        g = np.load(os.path.join(HERE, 'golden', f'bench_{name}.npz'))
        kw, X, y, th = getattr(bench, f'workload_{name}')()
        assert [zlib.crc32(np.ascontiguousarray(X).view(np.uint8)), zlib.crc32(np.ascontiguousarray(y).view(np.uint8))] == list(g['data_sum'])
        thetas = bench.theta_cloud(th, B, seed=seed)
        assert np.array_equal(thetas[g['rows']], g['thetas'])
        r = go.loglik(bench.oracle_spec(name), g['thetas'][2], X, y, want_grad=(name == 'c3'))
        assert abs(r.ll - g['ll'][2]) <= 1e-12 * abs(g['ll'][2])
        if name == 'c3':
            assert np.allclose(r.grad, g['grad'][2], rtol=1e-10, atol=1e-10 * np.max(np.abs(r.grad)))

"""Generates the committed golden fixtures.  Run in the BUILD container only:

    python tests/golden/make_golden.py

(a) transforms_ref.npz -- outputs of the REFERENCE's own ``andvaranaut/transform.py`` (imported by
    file path from /root/reference with ``pytensor`` stubbed; its NumPy ``con/rev/der`` paths need only
    numpy/scipy/sklearn), on seeded inputs.  These pin ``oracle/warp_oracle.py`` and the product's
    ``andvaranaut_b200/transform.py`` to the reference's code.
(b) tutorial_kat.npz   -- the known answers recorded in ``tutorial/tutorial.ipynb:362-369``.
(c) gp_oracle_*.npz    -- seeded GP cases (inputs, theta, ll, grad, mu, var) produced by the oracle
    itself ("parity unpinned": no reference GP output exists); they guard the oracle against drift
    and give the GPU tests fixed inputs.
/root/reference is not available on the GPU box; nothing at test time reads it.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference_transform():
    pt = types.ModuleType('pytensor')
    ptt = types.ModuleType('pytensor.tensor')
    pt.tensor = ptt
    pt.shared = lambda x: x
    sys.modules['pytensor'] = pt
    sys.modules['pytensor.tensor'] = ptt
    spec = importlib.util.spec_from_file_location('ref_transform', '/root/reference/andvaranaut/transform.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


WGP_CASES = [
    # name, stages, params, data kind, xdist interval
    ('c2_x', ['uniform', 'kumaraswamy'], [1.3, 0.8], 'unit', (0.0, 1.0)),
    ('c2_y', ['logarithm', 'sal', 'meanstd'], [0.1, 1.2, -0.3, 0.9], 'pos', None),
    ('affine_arcsinh', ['affine', 'arcsinh'], [0.2, 1.5, 0.1, 0.7, -0.2, 1.1], 'real', None),
    ('boxcox_meanstd', ['boxcox', 'meanstd'], [0.15], 'pos', None),
    ('sinharcsinh_stddev', ['sinharcsinh', 'stddev'], [0.2, 1.1], 'real', None),
    ('stdshift_pzero', ['stdshift', 'sal', 'pzero'], [0.3, -0.1, 0.9, 0.2, 1.3], 'real', None),
    ('minshift_log', ['minshift', 'affine'], [0.5, 2.0], 'real', None),
    ('maxmin_kuma', ['maxmin', 'kumaraswamy'], [0.7, 1.4], 'real', None),
    ('kuma_maxmin', ['kumaraswamy', 'maxmin'], [1.2, 0.9], 'unit', None),
]


def data_of(kind, rng, n=64):
    if kind == 'unit':
        return rng.uniform(0.02, 0.98, n)
    if kind == 'pos':
        return np.exp(rng.normal(0.0, 0.6, n))
    return rng.normal(0.3, 1.2, n)


def make_transforms(ref):
    rng = np.random.default_rng(20261018)
    out = {}
    x = rng.uniform(0.0, 2.0, 32)
    x2 = rng.uniform(1.0, 1.5, 32)
    yv = rng.normal(1.0, 2.0, 32)
    out['x_u'] = x
    out['x_n'] = x2
    out['y'] = yv
    u = ref.uniform(st.uniform(0, 2))
    out['uniform_con'] = u.con(x)
    out['uniform_rev'] = u.rev(u.con(x))
    nrm = ref.normal(st.uniform(1, 0.5))
    out['normal_con'] = nrm.con(x2)
    out['normal_rev'] = nrm.rev(nrm.con(x2))
    mm = ref.maxmin(x)
    out['maxmin_con'] = mm.con(x)
    out['maxmin_ab'] = np.array([mm.a, mm.b])
    mmc = ref.maxmin(x, centred=True)
    out['maxminc_con'] = mmc.con(x)
    ms = ref.meanstd(yv)
    out['meanstd_con'] = ms.con(yv)
    out['meanstd_ab'] = np.array([ms.a, ms.b])
    out['meanstd_rev'] = ms.rev(ms.con(yv))
    for nm, cls in [('probit', ref.probit), ('cdf', ref.cdf), ('logit_logistic', ref.logit_logistic)]:
        c = cls(st.norm(1.0, 2.0))
        out[nm + '_con'] = c.con(yv)
        out[nm + '_rev'] = c.rev(c.con(yv))
    ypos = np.exp(yv / 3)
    out['ypos'] = ypos
    for nm, cls in [('nonneg', ref.nonneg), ('log1p', ref.log1p), ('log10', ref.log10)]:
        c = cls()
        out[nm + '_con'] = c.con(ypos)
        out[nm + '_rev'] = c.rev(c.con(ypos))
    c = ref.normalise(3.5)
    out['normalise_con'] = c.con(yv)
    # composite warps, NumPy mode
    for name, stages, params, kind, interval in WGP_CASES:
        d = data_of(kind, rng)
        xd = st.uniform(interval[0], interval[1] - interval[0]) if interval is not None else None
        w = ref.wgp(stages, np.array(params, dtype=np.float64), y=d, xdist=xd)
        z = w.con(d)
        out[f'wgp_{name}_data'] = d
        out[f'wgp_{name}_con'] = z
        out[f'wgp_{name}_der'] = w.der(d)
        out[f'wgp_{name}_rev'] = w.rev(z)
        out[f'wgp_{name}_pos'] = w.pos.astype(np.int8)
        out[f'wgp_{name}_pid'] = w.pid
        out[f'wgp_{name}_np'] = np.array(w.np)
        # fresh points pushed through the frozen statistics
        t = data_of(kind, rng, 16)
        out[f'wgp_{name}_test'] = t
        out[f'wgp_{name}_test_con'] = w.con(t)
        out[f'wgp_{name}_test_der'] = w.der(t)
    np.savez(os.path.join(HERE, 'transforms_ref.npz'), **out)
    print('transforms_ref.npz:', len(out), 'arrays')


def make_tutorial_kat(ref):
    x = np.array([[1.85531589, 1.24150338], [0.88811964, 1.28931285]])
    xc = np.array([[0.92765794, -0.05886629], [0.44405982, 0.2723674]])
    y = np.array([-0.0312707, -0.28639611])
    space = [st.uniform(loc=0, scale=2), st.uniform(loc=1, scale=0.5)]
    got = np.c_[ref.uniform(space[0]).con(x[:, 0]), ref.normal(space[1]).con(x[:, 1])]
    assert np.allclose(got, xc, atol=5e-9), got
    yt = np.array([xx[0] ** 2 - xx[0] - xx[1] ** 2 * xx[0] + xx[1] for xx in x])
    assert np.allclose(yt, y, atol=5e-9)
    np.savez(os.path.join(HERE, 'tutorial_kat.npz'), x=x, xc=xc, y=y)
    print('tutorial_kat.npz ok')


def gp_cases():
    from oracle.gp_oracle import ModelSpec
    return {
        'rbf_c1': dict(spec=ModelSpec(nx=2, kerns=['RBF'], noise=False), N=100, M=64, seed=101, tutorial=True),
        'm52_noise': dict(spec=ModelSpec(nx=4, kerns=['Matern52'], noise=True), N=96, M=50, seed=7),
        'm32_exp_sum': dict(spec=ModelSpec(nx=3, kerns=['Matern32', 'Exponential'], ops=['+'], noise=True), N=70, M=33, seed=8),
        'rbf_rq_prod': dict(spec=ModelSpec(nx=3, kerns=['RBF', 'RatQuad'], ops=['*'], noise=True), N=65, M=20, seed=9),
        'c2_small': dict(spec=ModelSpec(nx=3, kerns=['Matern52'], noise=True,
                                        xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 3,
                                        ywarp=['logarithm', 'sal', 'meanstd']), N=80, M=0, seed=202),
    }


def gp_inputs(case):
    from oracle.gp_oracle import ModelSpec
    spec, N, M, seed = case['spec'], case['N'], case['M'], case['seed']
    rng = np.random.default_rng(seed)
    d = spec.nx
    if case.get('tutorial'):
        # config C1: tutorial target fn and priors (tutorial.ipynb:61-68), maxmin/meanstd conversion
        # (:439-441), theta = the hypers the tutorial records (:529)
        P = st.qmc.LatinHypercube(d=2, seed=seed).random(N)
        xr = np.c_[2.0 * P[:, 0], 1.0 + 0.5 * P[:, 1]]
        yr = xr[:, 0] ** 2 - xr[:, 0] - xr[:, 1] ** 2 * xr[:, 0] + xr[:, 1]
        X = np.empty_like(xr)
        for j in range(2):
            xm = (xr[:, j].max() - xr[:, j].min()) / (1 - 2 * 0.01)
            X[:, j] = -xr[:, j].min() / xm + 0.01 + xr[:, j] / xm
        y = (yr - yr.mean()) / yr.std()
        th = np.array([1.1314017, 2.68928595, 68.35800214])
        Xs = rng.uniform(0.01, 0.99, (M, 2))
        return X, y, th, Xs
    X = st.qmc.LatinHypercube(d=d, seed=seed).random(N)
    a = np.linspace(0.5, 2.0, d)
    y = np.exp(np.sum(np.sin(2 * np.pi * a * X), axis=1) / d + 0.5 * X[:, 0] * X[:, -1]) + 0.01 * rng.normal(size=N)
    if spec.ywarp is None:
        y = (y - y.mean()) / y.std()
    o = spec.offsets()
    th = np.zeros(o['P'])
    if spec.noise:
        th[o['gv']] = 1e-3 * np.exp(0.3 * rng.normal())
    th[o['l']:o['l'] + d * spec.nkern] = np.exp(0.3 * rng.normal(size=d * spec.nkern))
    th[o['kv']:o['kv'] + spec.nkern] = 1.5 * np.exp(0.3 * rng.normal(size=spec.nkern))
    th[o['iw']:o['iw'] + spec.n_iw()] = np.exp(0.2 * rng.normal(size=spec.n_iw()))
    if spec.ywarp is not None:
        th[o['cw']:o['cw'] + 4] = [0.1, 1.1, -0.2, 0.9]
    if spec.has_alpha:
        th[o['alpha']] = 1.7
    Xs = rng.uniform(0, 1, (M, d)) if M else np.zeros((0, d))
    return X, y, th, Xs


def next_row_outputs(spec, th, Xc, z, Xs):
    """golden values of the "next" rows: the BO refine graph (latent mean / variance without the noise term
    and their gradients w.r.t. the query points, gpmcmc.py:738-778) at the first 8 test points, and the inverse-problem
    potential (gpmcmc.py:1098-1165, dense stacked system) at the first 4 for one observation with and without a
    variance.  ``inv_noise_t`` is what the reference puts on the training diagonal: sqrt(gv + jitter)."""
    from oracle import gp_oracle as go
    pm, pv, pdm, pdv = go.predict_grad(spec, th, Xc, z, Xs[:8], pred_noise=False)
    gv = go.unpack(spec, th)['gv']
    noise_t = np.sqrt(gv + spec.jitter)
    yo = np.array([0.3])
    out = dict(pg_mu=pm, pg_var=pv, pg_dmu=pdm, pg_dvar=pdv, inv_yo=yo, inv_noise_t=np.array(noise_t))
    for tag, noise_o in (('inv_ll_exact', 0.0), ('inv_ll_noisy', 0.05)):
        ynoise = np.r_[np.full(len(z), noise_t), noise_o]
        out[tag] = np.array([go.inverse_loglik(spec, th, Xc, z, x, yo, ynoise) for x in Xs[:4]])
    return out


def make_gp():
    from oracle import gp_oracle as go
    for name, case in gp_cases().items():
        X, y, th, Xs = gp_inputs(case)
        spec = case['spec']
        r = go.loglik(spec, th, X, y, want_grad=True, keep=True)
        out = dict(X=X, y=y, theta=th, ll=np.array(r.ll), grad=r.grad)
        if case['M']:
            mu, var = go.predict(spec, th, r.Xw, r.z, Xs)
            out.update(Xs=Xs, mu=mu, var=var)
            out.update(next_row_outputs(spec, th, r.Xw, r.z, Xs))
        np.savez(os.path.join(HERE, f'gp_oracle_{name}.npz'), **out)
        print(name, 'll', r.ll, 'P', len(th))


if __name__ == '__main__':
    ref = load_reference_transform()
    make_transforms(ref)
    make_tutorial_kat(ref)
    make_gp()

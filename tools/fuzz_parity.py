"""Randomised parity sweep of the CUDA path against the oracle (development aid; the committed tests hold the fixed
cases): random N, d, kernel folds, noise, batch sizes, test-batch sizes, appends.
    python tools/fuzz_parity.py [ncases] [seed]
Prints one line per failure and a summary; exit code 1 if anything exceeded the tolerances of tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import gp_oracle as go  # noqa: E402
import cases  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

KERNS = ['RBF', 'Matern52', 'Matern32', 'Exponential', 'RatQuad']


def rand_spec(rng):
    nk = int(rng.choice([1, 1, 1, 2, 3]))
    kerns = []
    for _ in range(nk):
        k = str(rng.choice(KERNS))
        if k == 'RatQuad' and 'RatQuad' in kerns:
            k = 'RBF'
        kerns.append(k)
    ops = [str(rng.choice(['+', '*'])) for _ in range(nk - 1)]
    return go.ModelSpec(nx=int(rng.integers(1, 13)), kerns=kerns, ops=ops, noise=bool(rng.random() < 0.8),
                        jitter=float(rng.choice([1e-6, 1e-4])))


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    worst = dict(ll=0.0, grad=0.0, mu=0.0, var=0.0, dmu=0.0, app=0.0)
    for ci in range(ncases):
        spec = rand_spec(rng)
        N = int(rng.choice([2, 3, 5, 17, 63, 64, 65, 100, 127, 128, 129, 200, 257, 300]))
        M = int(rng.choice([1, 5, 63, 64, 65, 300, 1000, 4097]))
        B = int(rng.choice([1, 2, 5]))
        X, y, th, Xs = cases.synth(spec, N, seed=int(rng.integers(1 << 30)), M=M)
        expo = 'Exponential' in spec.kerns
        try:
            eng = GPEngine(**cases.engine_args(spec))
            eng.set_data(X, y)
            ths = np.stack([th * np.exp(0.05 * rng.normal(size=th.shape)) for _ in range(B)])
            ll, gr, info = (t.cpu().numpy() for t in eng.loglik_grad(ths))
            for b in range(B):
                r = go.loglik(spec, ths[b], X, y, keep=True)
                if r.info != 0 or info[b] != 0:
                    assert (r.info != 0) == (info[b] != 0), ('info', r.info, info[b])
                    continue
                cond = np.linalg.cond(r.L) ** 2
                tol = max(1e-9, 50 * cond * np.finfo(float).eps)
                e_ll = abs(ll[b] - r.ll) / max(abs(r.ll), 1.0)
                e_g = np.max(np.abs(gr[b] - r.grad) / np.maximum(np.abs(r.grad), 1e-3 * np.max(np.abs(r.grad)) + 1e-300))
                worst['ll'] = max(worst['ll'], e_ll / tol)
                worst['grad'] = max(worst['grad'], e_g / (tol * (1e3 if expo else 1.0)))
                assert e_ll <= tol, ('ll', e_ll, tol)
                assert e_g <= tol * (1e3 if expo else 1.0), ('grad', e_g, tol)
            r0 = go.loglik(spec, th, X, y, keep=True)
            if r0.info == 0 and int(eng.factorize(th)[0]) == 0:
                cond = np.linalg.cond(r0.L) ** 2
                tol = max(1e-8, 50 * cond * np.finfo(float).eps)
                kv = go.kdiag_total(spec, go.unpack(spec, th)['kv'])
                mu_r, var_r = go.predict(spec, th, X, y, Xs)
                mu, var = (t.cpu().numpy() for t in eng.predict(Xs))
                e_mu = np.max(np.abs(mu - mu_r)) / max(np.max(np.abs(mu_r)), 1e-300)
                e_var = np.max(np.abs(var - var_r) / np.maximum(np.abs(var_r), kv))
                worst['mu'] = max(worst['mu'], e_mu / tol)
                worst['var'] = max(worst['var'], e_var / tol)
                assert e_mu <= tol and e_var <= tol, ('predict', e_mu, e_var, tol)
                Mg = min(M, 200)
                rm, rv, rdm, rdv = go.predict_grad(spec, th, X, y, Xs[:Mg], pred_noise=False)
                m, v, dm, dv = (t.cpu().numpy() for t in eng.predict_grad(Xs[:Mg], pred_noise=False))
                e_dm = np.max(np.abs(dm - rdm)) / max(np.max(np.abs(rdm)), 1e-300)
                e_dv = np.max(np.abs(dv - rdv)) / max(np.max(np.abs(rdv)), kv)
                worst['dmu'] = max(worst['dmu'], e_dm / tol)
                assert e_dm <= tol and e_dv <= tol * (1e3 if expo else 1.0), ('predict_grad', e_dm, e_dv, tol)
                # append 3 points, compare with the oracle on the enlarged set
                Xa, ya, _, _ = cases.synth(spec, 3, seed=int(rng.integers(1 << 30)))
                ok = True
                for q in range(3):
                    ok = ok and int(eng.append(Xa[q], ya[q])[0]) == 0
                if ok:
                    X2, y2 = np.r_[X, Xa], np.r_[y, ya]
                    r2 = go.loglik(spec, th, X2, y2, keep=True)
                    if r2.info == 0:
                        tol2 = max(1e-8, 50 * np.linalg.cond(r2.L) ** 2 * np.finfo(float).eps)
                        mu_r, var_r = go.predict(spec, th, X2, y2, Xs)
                        mu, var = (t.cpu().numpy() for t in eng.predict(Xs))
                        e_a = max(np.max(np.abs(mu - mu_r)) / max(np.max(np.abs(mu_r)), 1e-300),
                                  np.max(np.abs(var - var_r) / np.maximum(np.abs(var_r), kv)))
                        worst['app'] = max(worst['app'], e_a / tol2)
                        assert e_a <= tol2, ('append', e_a, tol2)
        except AssertionError as e:
            fails += 1
            print(f'FAIL case {ci}: N={N} d={spec.nx} kerns={spec.kerns} ops={spec.ops} noise={spec.noise} B={B} M={M}: {e}', flush=True)
        torch.cuda.synchronize()
    print(f'{ncases} cases, {fails} failures; worst error / tolerance:', {k: round(v, 3) for k, v in worst.items()})
    sys.exit(1 if fails else 0)


if __name__ == '__main__':
    main()

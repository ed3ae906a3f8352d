"""Generates tests/golden/gp_truth_*.npz: 50-digit evaluations (oracle/gp_truth.py, mpmath) of the GP arithmetic on the
float64 inputs of two cases, committed so that the tests can MEASURE rounding errors instead of assuming a floor:

  c1        the reference's tutorial configuration (C1: RBF, d=2, N=100, noise=False, jitter 1e-6, cond(K) ~ 6e9) at the
            hyperparameters the tutorial records (tutorial.ipynb:529) and 7 more around them
  m52       a well-conditioned Matern-5/2 case with noise (the inputs of gp_oracle_m52_noise.npz)

Run in the build container (takes a few minutes):   python tests/golden/make_truth.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from oracle import gp_truth  # noqa: E402


def main():
    cases = mg.gp_cases()
    for tag, name, nth, npred in (('c1', 'rbf_c1', 8, 2), ('m52', 'm52_noise', 3, 1)):
        case = cases[name]
        spec = case['spec']
        X, y, th, Xs = mg.gp_inputs(case)
        rng = np.random.default_rng(1000 + len(name))
        thetas = np.vstack([th[None, :], th[None, :] * np.exp(0.15 * rng.normal(size=(nth - 1, len(th))))])
        out = dict(X=X, y=y, thetas=thetas, Xs=Xs, ll=[], grad=[], cond_est=[], mu=[], var=[])
        for b, t in enumerate(thetas):
            t0 = time.time()
            r = gp_truth.evaluate(spec.kerns[0], spec.noise, spec.jitter, t, X, y, Xs if b < npred else None)
            out['ll'].append(r['ll'])
            out['grad'].append(r['grad'])
            out['cond_est'].append(r['cond_est'])
            if b < npred:
                out['mu'].append(r['mu'])
                out['var'].append(r['var'])
            print(tag, b, 'll', r['ll'], 'cond~', f"{r['cond_est']:.2e}", f'{time.time() - t0:.1f}s', flush=True)
        np.savez(os.path.join(HERE, f'gp_truth_{tag}.npz'), **{k: np.asarray(v) for k, v in out.items()})


if __name__ == '__main__':
    main()

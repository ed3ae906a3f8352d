"""C1 (BASELINE.json configs[0]): the tutorial workflow through the GPMCMC API -- LHC N=100, d=2, RBF, noise=False,
maxmin / meanstd conversions, MAP fit from the default start, predictions -- timed next to the wall-clock prints the
reference's tutorial notebook records (unknown CPU, PyMC graph compile + refactorisation inside every
predict call).    python tools/c1_tutorial_probe.py"""
import json
import os
import sys
import time

import numpy as np
import scipy.stats as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from andvaranaut_b200 import GPMCMC, maxmin, meanstd  # noqa: E402


def target(x):
    x1, x2 = x
    return np.array([x1 ** 2 - x1 - x2 ** 2 * x1 + x2])          # tutorial.ipynb:61-63


def run():
    import torch
    priors = [st.uniform(loc=0, scale=2), st.uniform(loc=1, scale=0.5)]
    g = GPMCMC(kernel='RBF', noise=False, nx=2, ny=1, priors=priors, target=target, verbose=False)
    g.sample(100, seed=101)
    g.change_conrevs([maxmin(g.x[:, 0]), maxmin(g.x[:, 1])], [meanstd(g.y[:, 0])])
    g.fit(maxeval=3)                                              # warm-up (library load, allocations)
    g.predict(g.x[:4])

    def timed(fn, reps=5):
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best
    t0 = time.perf_counter()
    data = g.fit(return_data=True)
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    x1 = np.column_stack([rng.uniform(0, 2, 1), rng.uniform(1, 1.5, 1)])
    x1k = g._LHC__latin_sample(1000, seed=1)
    x10k = g._LHC__latin_sample(10000, seed=2)
    out = {
        'workload': 'c1 tutorial: N=100 d=2 RBF noise=False, MAP fit + predict through GPMCMC',
        'map_fit_s': t_fit, 'map_fit_evals': int(data['evals']), 'logp': float(data['logp']),
        'predict_1_point_s': timed(lambda: g.predict(x1)),
        'predict_1000_points_s': timed(lambda: g.predict(x1k)),
        'predict_10000_candidates_s': timed(lambda: g.predict(x10k, return_var=True)),
        'rmse_1000': float(np.sqrt(np.mean((g.predict(x1k)[:, 0] - np.array([target(x)[0] for x in x1k])) ** 2))),
        'reference_tutorial_recorded': {'predict_1_point_s': [0.38, 0.78], 'predict_1000_points_s': 1.00,
                                        'predict_10000_candidates_s': [0.48, 1.05], 'map_fit_evals': 19,
                                        'rmse_heldout': 1.44e-4,
                                        'source': 'BASELINE.md section 1: tutorial/tutorial.ipynb:488,566-569,789-790,958-959,1022-1023 '
                                                  '(unknown CPU; includes PyTensor compile and refactorisation per call)'},
    }
    return out


if __name__ == '__main__':
    print(json.dumps(run()))

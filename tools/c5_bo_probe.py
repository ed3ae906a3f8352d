"""C5 (BASELINE.json configs[4]): Bayesian-optimisation iterations through the GPMCMC API at several training-set
sizes: d = 12, 4096 LHC candidates per iteration, EI acquisition, warm-started MAP refit after every new point
(SURVEY 8d).  Prints one JSON line per size with iterations/s and the split acquisition / refit.
    python tools/c5_bo_probe.py [sizes...]      default 256 1024 4096"""
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
    return np.array([np.sum((x - 0.3) ** 2) + np.sin(5.0 * x[0])])


def run(N, iters=3, d=12, seed=505):
    import torch
    g = GPMCMC(kernel='Matern52', noise=True, nx=d, ny=1, priors=[st.uniform(0, 1)] * d, target=target,
               xconrevs=[maxmin(np.array([0.0, 1.0])) for _ in range(d)], yconrevs=[meanstd(np.array([0.0, 1.0]))],
               verbose=False)
    rng = np.random.default_rng(seed)
    x = st.qmc.LatinHypercube(d=d, seed=seed).random(N)
    y = np.array([target(xi) for xi in x])
    g.set_data(x, y)
    g.change_yconrevs([meanstd(y)])
    t0 = time.perf_counter()
    g.fit(method='map')
    torch.cuda.synchronize()
    t_fit0 = time.perf_counter() - t0
    t_acq = t_fit = 0.0
    evals = []
    y0 = g.y.min()
    for it in range(iters):
        t0 = time.perf_counter()
        xs = g._LHC__latin_sample(4096, seed=int(rng.integers(2 ** 31)))
        ei = g.predict(xs, EI=True, EIopt='min')[:, 0]
        xn = xs[int(np.argmax(ei))][None, :]
        torch.cuda.synchronize()
        t_acq += time.perf_counter() - t0
        t0 = time.perf_counter()
        g.set_data(np.r_[g.x, xn], np.r_[g.y, np.array([target(xn[0])])])
        data = g.fit(method='map', start=g.hypers, return_data=True)
        torch.cuda.synchronize()
        t_fit += time.perf_counter() - t0
        evals.append(int(data['evals']))
    return {'N': N, 'd': d, 'candidates': 4096, 'iters': iters, 'first_fit_s': round(t_fit0, 3),
            'acquire_ms': round(1e3 * t_acq / iters, 2), 'refit_ms': round(1e3 * t_fit / iters, 2),
            'iters_per_s': round(iters / (t_acq + t_fit), 3), 'best_y': float(g.y.min()), 'start_best_y': float(y0),
            'refit_evals': evals}


if __name__ == '__main__':
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]
    for n in sizes:
        print(json.dumps(run(n)), flush=True)

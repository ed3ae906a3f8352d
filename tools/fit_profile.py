"""cProfile of the C2 MAP fit through GPMCMC.fit (development aid): where the host time of one optimiser step goes.
    python tools/fit_profile.py"""
import cProfile
import os
import pstats
import sys

import numpy as np
import scipy.stats as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200 import GPMCMC, wgp  # noqa: E402

kw, X, y, th = bench.workload_c2()
N, d = X.shape
priors = [st.uniform(0, 1)] * d


def build():
    g = GPMCMC(kernel='Matern52', noise=True, nx=d, ny=1, priors=priors, target=lambda x: np.zeros(1), verbose=False,
               xconrevs=[wgp(['uniform', 'kumaraswamy'], [1.0, 1.0], xdist=priors[i]) for i in range(d)],
               yconrevs=[wgp(['logarithm', 'sal', 'meanstd'], [0.0, 1.0, 0.0, 1.0], y=y)])
    g.set_data(X, y[:, None])
    return g


g = build()
g.fit(iwgp=True, cwgp=True, maxeval=3)
g = build()
pr = cProfile.Profile()
pr.enable()
data = g.fit(iwgp=True, cwgp=True, return_data=True)
pr.disable()
print('evals', data['evals'])
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

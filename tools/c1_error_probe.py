"""Where the device's rounding error on the ill-conditioned tutorial case (C1) comes from: per hyperparameter vector of
tests/golden/gp_truth_c1.npz, the relative ll error against the 50-digit truth of (a) the NumPy oracle, (b) the device,
(c) LAPACK Cholesky on the DEVICE's covariance matrix, (d) LAPACK Cholesky on the correctly rounded covariance matrix;
and the distance of the oracle's / device's K from the correctly rounded one in ulps.    python tools/c1_error_probe.py"""
import os
import sys

import mpmath as mp
import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden as mg  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

g = np.load(os.path.join(ROOT, 'tests', 'golden', 'gp_truth_c1.npz'))
spec = mg.gp_cases()['rbf_c1']['spec']
X, y, thetas = g['X'], g['y'], g['thetas']
N = len(y)
eng = GPEngine(nx=2, kerns=['RBF'], noise=False, jitter=spec.jitter)
eng.set_data(X, y)
ll_dev = eng.loglik_grad(thetas)[0].cpu().numpy()
K_dev = eng.cov(thetas).cpu().numpy()[:, :N, :N]


def ll_of(K):
    L = sla.cholesky(K, lower=True)
    b = sla.solve_triangular(L, y, lower=True)
    return -0.5 * N * np.log(2 * np.pi) - 0.5 * b @ b - np.sum(np.log(np.diag(L)))


mp.mp.dps = 40
for b, t in enumerate(thetas):
    th = go.unpack(spec, t)
    Ko = go.cov_matrix(spec, th, X) + spec.jitter * np.eye(N)
    l = [mp.mpf(float(v)) for v in th['l']]
    kv = mp.mpf(float(th['kv'][0]))
    Kx = np.empty((N, N))
    for i in range(N):
        for j in range(i + 1):
            r2 = sum(((mp.mpf(float(X[i, m])) - mp.mpf(float(X[j, m]))) / l[m]) ** 2 for m in range(2))
            Kx[i, j] = Kx[j, i] = float(kv * mp.exp(-r2 / 2) + (mp.mpf(float(spec.jitter)) if i == j else 0))
    Kd = np.tril(K_dev[b]) + np.tril(K_dev[b], -1).T
    tr = g['ll'][b]
    ulp = np.spacing(Kx)
    print(f'theta {b}: ll err oracle {(go.loglik(spec, t, X, y, want_grad=False).ll - tr) / abs(tr):+.2e}  device {(ll_dev[b] - tr) / abs(tr):+.2e}  '
          f'lapack(device K) {(ll_of(Kd) - tr) / abs(tr):+.2e}  lapack(exact K) {(ll_of(Kx) - tr) / abs(tr):+.2e}   '
          f'K vs exact, ulps: oracle max {np.max(np.abs(Ko - Kx) / ulp):.1f} mean {np.mean((Ko - Kx) / ulp):+.3f}; '
          f'device max {np.max(np.abs(Kd - Kx) / ulp):.1f} mean {np.mean((Kd - Kx) / ulp):+.3f}', flush=True)

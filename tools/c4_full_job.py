"""C4 (BASELINE.json configs[3]) as ONE job through the engine API: predictive mean / variance with Gauss-Hermite
reversion for 10^7 test points from an N=8192, d=10 Matern-5/2 GP.  Host buffers in, host buffers out: the timed region
holds the factorisation, the host->device copy of the test points (pinned, in chunks), every predict launch and the
device->host copy of mean and variance.  Prints one JSON line.
    python tools/c4_full_job.py [M]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
kw, X, y, th = bench.workload_c4()
dev = torch.device('cuda:0')
eng = GPEngine(**kw, device=dev)
eng.set_data(X, y)
epi = GPEngine.make_epilogue(mode='revert', deg=8, yrev=[(0, -1, (0.0, 1.0, 0.0, 0.0))])
chunk = 148 * 128 * 8                                   # 151552 points per host chunk
xs_host = torch.from_numpy(np.random.default_rng(405).uniform(size=(M, 10))).pin_memory()
mu_host = torch.empty(M, dtype=torch.float64).pin_memory()
var_host = torch.empty(M, dtype=torch.float64).pin_memory()
eng.factorize(th)
eng.predict(xs_host[:chunk].to(dev), epilogue=epi)      # warm-up: allocations, smem opt-in
torch.cuda.synchronize()
t0 = time.perf_counter()
info = eng.factorize(th)
launches = 0
for s in range(0, M, chunk):
    xd = xs_host[s:s + chunk].to(dev, non_blocking=True)
    mu, var = eng.predict(xd, epilogue=epi)
    launches += eng.launches
    mu_host[s:s + chunk].copy_(mu, non_blocking=True)
    var_host[s:s + chunk].copy_(var, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
N = X.shape[0]
print(json.dumps({'workload': f'c4 full job: N={N} d=10 Matern52, {M} test points, mean+variance+GH(8) reversion',
                  'seconds': dt, 'points_per_s': M / dt, 'launches': launches, 'info': int(info[0]),
                  'h2d_bytes': M * 10 * 8, 'd2h_bytes': M * 16,
                  'frac_of_fp64_peak_at_35.5TF': M * bench.flops_predict(N, 10) / dt / 35.5e12,
                  'mean_range': [float(mu_host.min()), float(mu_host.max())],
                  'var_range': [float(var_host.min()), float(var_host.max())]}))

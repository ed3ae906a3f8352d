"""Which column bounds warp_kernel for one sample of the c2 model: only the input warps, only the output warp, both
(development aid).   python tools/warp_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

kw, X, y, th = bench.workload_c2()
d = X.shape[1]
ylog = (np.log(y) - np.log(y).mean()) / np.log(y).std()
for name, kws, yy, theta in (
        ('both', kw, y, th),
        ('x warps only', dict(kw, ywarp=None), ylog, th[:-4]),
        ('y warp only', dict(kw, xwarps=None), y, np.concatenate([th[:d + 2], th[-4:]]))):
    eng = GPEngine(**kws, device='cuda:0')
    eng.set_data(X, yy)
    t = torch.as_tensor(theta[None, :], device='cuda:0')
    for _ in range(3):
        eng.loglik_grad(t)
    eng.set_profiling(True)
    acc = {}
    for _ in range(5):
        eng.loglik_grad(t)
        eng.loglik_grad(t)
        for k, v in eng.phase_ms().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    print(name, {k: round(v, 4) for k, v in acc.items() if v > 0})

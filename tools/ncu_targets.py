"""Small programs to run under ncu (development aid): each runs one warm-up pass and one measured pass of a bench
workload so that `ncu -k regex:... --launch-skip <matched launches of the warm-up> -c <n>` captures the second pass.
    python tools/ncu_targets.py c2 64      # one batched loglik+grad call of the c2 workload, B=64
    python tools/ncu_targets.py c3 512
    python tools/ncu_targets.py c4         # factorize + one K_xs panel (18944 points) of the c4 predict, GH epilogue
    python tools/ncu_targets.py c2b1       # single evaluation (B=1) of the c2 workload"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200 import transform as T  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
dev = 'cuda:0'
if wl in ('c2', 'c3', 'c2b1'):
    name = 'c2' if wl == 'c2b1' else wl
    B = 1 if wl == 'c2b1' else (int(sys.argv[2]) if len(sys.argv) > 2 else 64)
    kw, X, y, th = getattr(bench, 'workload_' + name)()
    eng = GPEngine(**kw, device=dev)
    eng.set_data(X, y)
    thetas = torch.as_tensor(bench.theta_cloud(th, B, seed=202 if name == 'c2' else 303), device=dev)
    for _ in range(2):
        ll, g, info = eng.loglik_grad(thetas)
    torch.cuda.synchronize()
    print(wl, 'B', B, 'll[0]', float(ll[0]), 'info', int(info.abs().sum()))
else:
    (kw, X, y, th), ab = bench.workload_c4()
    eng = GPEngine(**kw, device=dev)
    eng.set_data(X, y)
    info = eng.factorize(th)
    print('factorize info', int(info[0]))
    M = 18944
    Xs = torch.as_tensor(bench.c4_test_points(M), device=dev)
    epi = GPEngine.make_epilogue(mode='revert', deg=8, yrev=[(T.OP_AFFINE_CONST, -1, (ab[0], ab[1], 0.0, 0.0))])
    for _ in range(2):
        mu, var = eng.predict(Xs, epilogue=epi)
    torch.cuda.synchronize()
    print('c4 M', M, 'mu[0]', float(mu[0]), 'var[0]', float(var[0]))

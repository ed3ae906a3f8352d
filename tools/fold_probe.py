"""One batched loglik+grad call of a two-kernel fold at the C2 shape (for an ncu capture of kinv_grad_fold2_kernel).
    python tools/fold_probe.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
_, X, y, _ = bench.workload_c2()
y = (np.log(y) - np.log(y).mean()) / np.log(y).std()
d = 8
eng = GPEngine(nx=d, kerns=['RBF', 'Matern52'], ops=['*'], noise=True, device='cuda:0')
eng.set_data(X, y)
th = np.r_[1e-4, 0.7 * np.ones(2 * d), 1.5 * np.ones(2)]
ths = torch.as_tensor(bench.theta_cloud(th, B, seed=1), device='cuda:0')
for _ in range(2):
    ll, g, info = eng.loglik_grad(ths)
torch.cuda.synchronize()
print(float(ll[0]), int(info.sum()))

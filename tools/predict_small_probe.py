"""Device time of predict / predict_grad for SMALL test batches (BO candidates, refine and inverse-problem starts):
    python tools/predict_small_probe.py
One JSON line per (N, M): milliseconds per call (CUDA events, 5 calls after 2 warm-up calls) and launches."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from andvaranaut_b200.gp import GPEngine  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    d = 12
    rng = np.random.default_rng(0)
    for N in (1024, 4096, 8100):
        X = rng.uniform(size=(N, d))
        y = np.sin(X @ np.linspace(0.5, 2.0, d)) + 0.01 * rng.normal(size=N)
        eng = GPEngine(nx=d, kerns=['Matern52'], noise=True, device='cuda:0')
        eng.set_data(X, y)
        eng.factorize(np.r_[1e-4, np.ones(d), 1.5])
        epi = GPEngine.make_epilogue(mode='EI', deg=8, EIopt='min', yopt=0.0, yrev=[(0, -1, (0.0, 1.0, 0.0, 0.0))])
        for M in (1, 64, 256, 4096, 16384):
            Xs = torch.as_tensor(rng.uniform(size=(M, d)), device='cuda:0')
            tp = timed(lambda: eng.predict(Xs, epilogue=epi))
            lp = eng.launches
            tg = timed(lambda: eng.predict_grad(Xs, epilogue=epi, pred_noise=False))
            print(json.dumps({'N': N, 'M': M, 'predict_ms': round(tp, 4), 'predict_launches': lp,
                              'predict_grad_ms': round(tg, 4), 'predict_grad_launches': eng.launches,
                              'predict_TF': round(M * N * N / tp / 1e9, 2)}), flush=True)


if __name__ == '__main__':
    main()

"""Timing of the c4 / c5 predict calls (development aid): python tools/predict_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200 import transform as T  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


(kw, X, y, th), ab = bench.workload_c4()
eng = GPEngine(**kw, device='cuda:0')
eng.set_data(X, y)
eng.factorize(th)
epi = GPEngine.make_epilogue(mode='revert', deg=8, yrev=[(T.OP_AFFINE_CONST, -1, (ab[0], ab[1], 0.0, 0.0))])
for M in (148 * 128 * 6, 4096, 256):
    Xs = torch.as_tensor(bench.c4_test_points(M), device='cuda:0')
    ms = timed(lambda: eng.predict(Xs, epilogue=epi), 3 if M > 10000 else 20)
    print(f'c4 N=8192 M={M}: {ms:.3f} ms, {M / ms * 1e3:.0f} pts/s, {M * bench.flops_predict(8192, 10) / ms / 1e9:.2f} TF')
(kw5, X5, y5, th5), ab5, yopt5 = bench.workload_c5()
eng5 = GPEngine(**kw5, device='cuda:0')
eng5.set_data(X5, y5)
eng5.factorize(th5)
cand = torch.as_tensor(bench.c5_candidates(), device='cuda:0')
epi5 = GPEngine.make_epilogue(mode='EI', deg=8, EIopt='min', yopt=yopt5, yrev=[(T.OP_AFFINE_CONST, -1, (ab5[0], ab5[1], 0.0, 0.0))])
ms = timed(lambda: eng5.predict(cand, epilogue=epi5), 20)
print(f'c5 N=4096 M=4096: {ms:.3f} ms, {4096 / ms * 1e3:.0f} pts/s')

"""Per-kernel timing of one batched loglik+grad call on a bench workload (development aid).
    python tools/phase_probe.py c3 512      # workload, batch
With AVN_GP_LIB=build/libavn_gp_prof.so (built with -DAVN_FACTOR_PROF) the factor kernel also prints where its
CTAs spend their cycles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
grad = (sys.argv[3] != 'nograd') if len(sys.argv) > 3 else True
kw, X, y, th = getattr(bench, 'workload_' + wl)()
eng = GPEngine(**kw, device='cuda:0')
eng.set_data(X, y)
thetas = torch.as_tensor(bench.theta_cloud(th, B, seed=1), device='cuda:0')
for _ in range(3):
    eng.loglik_grad(thetas, want_grad=grad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.loglik_grad(thetas, want_grad=grad)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
eng.set_profiling(True)
eng.loglik_grad(thetas, want_grad=grad)
pm = eng.phase_ms()
N, d = X.shape
print(f'{wl} B={B} N={N} grad={grad}: {ms:.3f} ms/call, {B / ms * 1e3:.1f} evals/s, '
      f'{B * bench.flops_ll(N, d) / ms / 1e9:.2f} TF algorithmic; phases',
      {k: round(v, 3) for k, v in pm.items() if v > 0})

"""Stage-by-stage comparison of the CUDA path with the oracle on the golden cases (diagnostics;
the judged parity tests live in tests/)."""
import os
import sys
import json

import numpy as np
import scipy.linalg as sla
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from oracle import gp_oracle as go  # noqa: E402
from andvaranaut_b200.gp import GPEngine  # noqa: E402
import cases  # noqa: E402
import make_golden as mg  # noqa: E402


def rel(a, b, floor=1e-300):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), floor))


def run_case(name, spec, X, y, th, Xs, B=3):
    N = X.shape[0]
    eng = GPEngine(**cases.engine_args(spec))
    eng.set_data(X, y)
    rng = np.random.default_rng(1)
    thetas = np.stack([th * np.exp(0.02 * rng.normal(size=th.shape)) if b else th for b in range(B)])
    ll, grad, info = eng.loglik_grad(thetas)
    torch.cuda.synchronize()
    bufs = eng.debug_buffers()
    out = {'case': name, 'N': N, 'P': len(th), 'launches': int(eng.launches), 'info': info.cpu().tolist()}
    for b in range(B):
        r = go.loglik(spec, thetas[b], X, y, want_grad=True, keep=True)
        K = go.cov_matrix(spec, go.unpack(spec, thetas[b]), r.Xw)
        K[np.diag_indices(N)] += go.unpack(spec, thetas[b])['gv'] + spec.jitter
        Kg = eng.cov(thetas[b])[0, :N, :N].cpu().numpy()
        Lg = bufs['kl'][b, :N, :N].cpu().numpy()
        Tg = bufs['t'][b, :N, :N].cpu().numpy()
        Tref = sla.solve_triangular(r.L, np.eye(N), lower=True)
        gnorm = np.abs(r.grad) + 1e-3 * np.max(np.abs(r.grad))
        o = dict(b=b,
                 xw=rel(bufs['xw'][b, :N].cpu().numpy(), r.Xw), z=rel(bufs['z'][b, :N].cpu().numpy(), r.z),
                 K=rel(np.tril(Kg), np.tril(K)), L=rel(np.tril(Lg), r.L), T=rel(np.tril(Tg), Tref),
                 beta=rel(bufs['beta'][b, :N].cpu().numpy(), r.beta), alpha=rel(bufs['alpha'][b, :N].cpu().numpy(), r.alpha),
                 ll_ref=r.ll, ll_gpu=float(ll[b]), ll_rel=abs(float(ll[b]) - r.ll) / abs(r.ll),
                 grad_rel=float(np.max(np.abs(grad[b].cpu().numpy() - r.grad) / gnorm)),
                 cond=float(np.linalg.cond(K)))
        out[f'b{b}'] = o
        if b == 0:
            out['grad_gpu'] = grad[b].cpu().numpy().tolist()
            out['grad_ref'] = r.grad.tolist()
    if Xs is not None and len(Xs) and spec.xwarps is None and spec.ywarp is None:
        eng.factorize(th)
        mu, var = eng.predict(Xs)
        mu_r, var_r = go.predict(spec, th, X, y, Xs)
        kv = go.kdiag_total(spec, go.unpack(spec, th)['kv'])
        out['pred'] = dict(mu=rel(mu.cpu().numpy(), mu_r),
                           var_mixed=float(np.max(np.abs(var.cpu().numpy() - var_r) / np.maximum(np.abs(var_r), kv))))
    return out


if __name__ == '__main__':
    res = []
    for name, case in mg.gp_cases().items():
        X, y, th, Xs = mg.gp_inputs(case)
        try:
            o = run_case(name, case['spec'], X, y, th, Xs)
        except Exception as e:  # keep going: one broken case must not hide the others
            o = {'case': name, 'error': repr(e)}
        res.append(o)
        print(json.dumps(o))
    # larger shapes
    for name, spec, N in [('c3_like', go.ModelSpec(nx=6, kerns=['RBF'], noise=True), 1000),
                          ('m52_700', go.ModelSpec(nx=8, kerns=['Matern52'], noise=True), 700)]:
        X, y, th, Xs = cases.synth(spec, N, seed=303, M=300)
        try:
            o = run_case(name, spec, X, y, th, Xs, B=2)
        except Exception as e:
            o = {'case': name, 'error': repr(e)}
        res.append(o)
        print(json.dumps(o))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'gpu_diag.json'), 'w') as f:
        json.dump(res, f, indent=1)

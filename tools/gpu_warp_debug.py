import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from oracle import gp_oracle as go
from andvaranaut_b200.gp import GPEngine
import cases
np.set_printoptions(precision=6, linewidth=200)
specs = {
 'xonly': go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 3),
 'x1only': go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, xwarps=[None, (['uniform', 'kumaraswamy'], (0.0, 1.0)), None]),
 'yonly': go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, ywarp=['logarithm', 'sal', 'meanstd']),
 'ylog': go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, ywarp=['logarithm']),
 'yaff': go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, ywarp=['affine']),
}
for name, spec in specs.items():
    X, y, th, _ = cases.synth(spec, 80, seed=5)
    eng = GPEngine(**cases.engine_args(spec))
    eng.set_data(X, y)
    ll, grad, info = eng.loglik_grad(th[None, :])
    torch.cuda.synchronize()
    b = eng.debug_buffers()
    r = go.loglik(spec, th, X, y, want_grad=True, keep=True)
    print('====', name, 'P', eng.P, 'll', float(ll[0]), r.ll, 'info', info.tolist())
    print('theta', th)
    xw = b['xw'][0, :80].cpu().numpy(); z = b['z'][0, :80].cpu().numpy()
    print('xw gpu', xw[:3].ravel()); print('xw ref', r.Xw[:3].ravel()); print('X raw', X[:3].ravel())
    print('z gpu', z[:5]); print('z ref', r.z[:5]); print('y raw', y[:5])
    print('wstat', b['wstat'][0].cpu().numpy())
    print('grad gpu', grad[0].cpu().numpy()); print('grad ref', r.grad)

import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from oracle import gp_oracle as go
import scipy.linalg as sla
from andvaranaut_b200.gp import GPEngine
import cases
np.set_printoptions(precision=5, linewidth=220)
spec = go.ModelSpec(nx=3, kerns=['Matern52'], noise=True, xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 3)
X, y, th, _ = cases.synth(spec, 80, seed=5)
eng = GPEngine(**cases.engine_args(spec)); eng.set_data(X, y)
ll, grad, info = eng.loglik_grad(th[None, :]); torch.cuda.synchronize()
b = eng.debug_buffers()
thd = go.unpack(spec, th)
Xw, dXw = go.warp_inputs(spec, thd, X, True)
r = go.loglik(spec, th, X, y, keep=True)
N = 80
dx = b['dxw'][0, :N].cpu().numpy()   # [N,d,8]
for m in range(3):
    print('dim', m, 'dual err', np.max(np.abs(dx[:, m, :2] - dXw[:, m, 2*m:2*m+2])), 'gpu', dx[:2, m, :3].ravel(), 'ref', dXw[:2, m, 2*m:2*m+2].ravel())
G = b['gxpart'][0, :, :N, :].cpu().numpy().sum(axis=0)
# oracle GX
K, coef, parts = go.cov_matrix(spec, thd, Xw, None, want_parts=True)
Kinv = sla.cho_solve((r.L, True), np.eye(N)); W = np.outer(r.alpha, r.alpha) - Kinv
r2, kk, dk = parts[0]; WK = W * thd['kv'][0] * dk
GX = np.zeros((N, 3))
for m in range(3):
    D = Xw[:, m][:, None] - Xw[:, m][None, :]
    GX[:, m] = 2 * np.sum(WK * D, axis=1) / thd['l'][m] ** 2
print('GX err per dim', np.max(np.abs(G - GX), axis=0), 'scale', np.max(np.abs(GX), axis=0))
print('G gpu', G[:3]); print('G ref', GX[:3])
print('grad gpu', grad[0].cpu().numpy()); print('grad ref', r.grad)
print('recomputed iw grad from gpu pieces', np.einsum('nm,nmp->mp', G, dx[:, :, :2]).ravel())

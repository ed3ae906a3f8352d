"""Test double for GPEngine: same Python surface, arithmetic by the CPU oracle (tests only -- lets the host-side
drivers, the GPMCMC class and the gloo sharding tests run without a GPU)."""
import numpy as np
import torch

from oracle import gp_oracle as go


class OracleEngine:
    device = 'cpu'

    def __init__(self, spec):
        self.spec = spec
        self.P = spec.offsets()['P']
        self.launches = 0
        self.calls = []

    def set_data(self, X, y):
        self.X, self.y = np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64).reshape(-1)

    def loglik_grad(self, theta, want_grad=True, out=None):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        self.calls.append(theta.shape[0])
        ll, g, info = [], [], []
        for t in theta:
            r = go.loglik(self.spec, t, self.X, self.y, want_grad=True)
            ll.append(r.ll)
            g.append(r.grad)
            info.append(r.info)
        return (torch.tensor(ll, dtype=torch.float64), torch.tensor(np.array(g), dtype=torch.float64),
                torch.tensor(info, dtype=torch.int32))

    def factorize(self, theta):
        self.theta = np.asarray(theta, dtype=np.float64)
        return torch.zeros(1, dtype=torch.int32)

    def predict(self, Xs, epilogue=None, mean_add=None, **kw):
        mu, var = go.predict(self.spec, self.theta, self.X, self.y, np.asarray(Xs, dtype=np.float64))
        return torch.tensor(mu), torch.tensor(var)

    def predict_grad(self, Xs, epilogue=None, mean_add=None, dmean_add=None, pred_noise=True, **kw):
        out = go.predict_grad(self.spec, self.theta, self.X, self.y, np.asarray(Xs, dtype=np.float64), pred_noise=pred_noise)
        return tuple(torch.tensor(a) for a in out)

"""Multi-GPU sharding of the independent units of the GP path (one process per GPU, torch.distributed).

The path has no data-path exchange: MCMC chains / hyperparameter samples / MAP restarts and test-point blocks
are independent, the training set (<= 0.7 MB) and, for predict, the factor T (<= 537 MB at N = 8192) are
replicated, and the only collective is an all_gather of the small per-shard results ([B/G, 1+P] likelihoods and
gradients, [M/G, 2] predictions) -- NCCL on GPU tensors, gloo on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


class Shard:
    def __init__(self, group=None):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    def bounds(self, n):
        """contiguous block [lo, hi) of n units owned by this rank; every rank's block has at most ``per`` units."""
        per = -(-n // self.world)
        lo = min(n, self.rank * per)
        return lo, min(n, lo + per), per

    def _gather_rows(self, local, per, n):
        """local [k, C] (k <= per) -> [n, C] on every rank."""
        C = local.shape[1]
        buf = torch.zeros(per, C, dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
        if self.world == 1:
            return buf[:n]
        out = torch.empty(self.world * per, C, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        return out[:n]

    def loglik_grad(self, engine, theta):
        """theta [B,P] (numpy, identical on every rank) -> (ll [B], grad [B,P], info [B]) numpy on every rank."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        B, P = theta.shape
        lo, hi, per = self.bounds(B)
        if hi > lo:
            ll, g, info = engine.loglik_grad(theta[lo:hi])
            packed = torch.cat([ll[:, None], g, info.to(torch.float64)[:, None]], dim=1)
        else:
            dev = getattr(engine, 'device', 'cpu')
            packed = torch.zeros(0, P + 2, dtype=torch.float64, device=dev)
        full = self._gather_rows(packed, per, B).cpu().numpy()
        info = full[:, 1 + P].astype(np.int32)
        if np.any(info < 0):       # a rank's factor kernel aborted (include/avn_gp.h): every rank sees it and raises
            raise RuntimeError('avn_gp_loglik_grad: factorisation aborted on the device (info = -1) on at least one rank')
        return full[:, 0], full[:, 1:1 + P].copy(), info

    def predict(self, engine, Xs, **kw):
        """Xs [M,d] (identical on every rank) -> (mean [M], var [M]) numpy on every rank; blocks of test points are
        sharded, the factorisation is replicated (engine.factorize must have been called on every rank)."""
        Xs = np.asarray(Xs, dtype=np.float64)
        M = Xs.shape[0]
        lo, hi, per = self.bounds(M)
        madd = kw.pop('mean_add', None)
        if hi > lo:
            mu, var = engine.predict(Xs[lo:hi], mean_add=None if madd is None else madd[lo:hi], **kw)
            packed = torch.stack([mu, var], dim=1)
        else:
            packed = torch.zeros(0, 2, dtype=torch.float64, device=getattr(engine, 'device', 'cpu'))
        full = self._gather_rows(packed, per, M).cpu().numpy()
        return full[:, 0].copy(), full[:, 1].copy()

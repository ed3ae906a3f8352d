"""Multi-GPU sharding of the independent units of the GP path (one process per GPU, torch.distributed).

The path has no data-path exchange: MCMC chains / hyperparameter samples / MAP restarts and test-point blocks
are independent, the training set (<= 0.7 MB) and, for predict, the factor T (<= 537 MB at N = 8192) are
replicated, and the only collective is an all_gather of the small per-shard results ([B/G, 2+P] likelihoods,
gradients and info, [M/G, 2] predictions) -- NCCL on GPU tensors, gloo on CPU tensors (tests).

Two levels: the ``*_dev`` methods take and return DEVICE tensors (the rank's own block in, the gathered result of
all ranks out; nothing touches the host); ``loglik_grad`` / ``predict`` are the host-array calls the drivers and
``GPMCMC`` make (identical NumPy input on every rank, identical NumPy output on every rank).
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

    def _gather_rows(self, local, per, n, out=None):
        """local [k, C] (k <= per) -> [n, C] on every rank (rank-major; ``out`` = reusable [world*per, C] buffer)."""
        C = local.shape[1]
        if self.world == 1 and local.shape[0] == n:
            return local
        if local.shape[0] == per:
            buf = local.contiguous()
        else:
            buf = torch.zeros(per, C, dtype=local.dtype, device=local.device)
            buf[:local.shape[0]] = local
        if self.world == 1:
            return buf[:n]
        if out is None:
            out = torch.empty(self.world * per, C, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, buf, group=self.group)
        return out[:n]

    # ---- device level ------------------------------------------------------------------------------------
    def loglik_grad_dev(self, engine, theta_local, B_total, packed=None, gathered=None):
        """theta_local: this rank's block of the [B_total, P] batch (device tensor or array, may be empty).  Returns
        the gathered [B_total, P + 2] tensor (columns: ll, grad[P], info) on every rank.  ``packed`` [per, P+2] and
        ``gathered`` [world*per, P+2] are optional reusable buffers."""
        P = engine.P
        lo, hi, per = self.bounds(B_total)
        k = hi - lo
        if packed is None:
            packed = torch.zeros(per, P + 2, dtype=torch.float64, device=getattr(engine, 'device', 'cpu'))
        if k > 0:
            ll, g, info = engine.loglik_grad(theta_local)
            packed[:k, 0] = ll
            packed[:k, 1:1 + P] = g
            packed[:k, 1 + P] = info
        return self._gather_rows(packed if k == per else packed[:k], per, B_total, out=gathered)

    def predict_dev(self, engine, Xs_local, M_total, gathered=None, **kw):
        """Xs_local: this rank's block of the [M_total, d] test points.  Returns the gathered [M_total, 2] tensor
        (mean, variance) on every rank; the factorisation is replicated (``engine.factorize`` on every rank first)."""
        lo, hi, per = self.bounds(M_total)
        if hi > lo:
            mu, var = engine.predict(Xs_local, **kw)
            packed = torch.stack([mu, var], dim=1)
        else:
            packed = torch.zeros(0, 2, dtype=torch.float64, device=getattr(engine, 'device', 'cpu'))
        return self._gather_rows(packed, per, M_total, out=gathered)

    # ---- host level (what drivers.Posterior and GPMCMC call) -----------------------------------------------
    def loglik_grad(self, engine, theta):
        """theta [B,P] (numpy, identical on every rank) -> (ll [B], grad [B,P], info [B]) numpy on every rank."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        B, P = theta.shape
        lo, hi, per = self.bounds(B)
        full = self.loglik_grad_dev(engine, theta[lo:hi], B).cpu().numpy()
        info = full[:, 1 + P].astype(np.int32)
        if np.any(info < 0):       # a rank's factor kernel aborted (include/avn_gp.h): every rank sees it and raises
            raise RuntimeError('avn_gp_loglik_grad: factorisation aborted on the device (info = -1) on at least one rank')
        return full[:, 0].copy(), full[:, 1:1 + P].copy(), info

    def predict(self, engine, Xs, **kw):
        """Xs [M,d] (identical on every rank) -> (mean [M], var [M]) numpy on every rank; blocks of test points are
        sharded, the factorisation is replicated (engine.factorize must have been called on every rank)."""
        Xs = np.asarray(Xs, dtype=np.float64)
        M = Xs.shape[0]
        lo, hi, per = self.bounds(M)
        madd = kw.pop('mean_add', None)
        full = self.predict_dev(engine, Xs[lo:hi], M, mean_add=None if madd is None else madd[lo:hi], **kw).cpu().numpy()
        return full[:, 0].copy(), full[:, 1].copy()

"""Hyperparameter space of the GP model: priors, constraining transforms and the mapping between the
named PyMC-style variables of ``GPMCMC.hypers`` and the flat constrained vector the CUDA path consumes.

Restates what ``GPMCMC.__fit`` declares inside ``pm.Model()`` (andvaranaut/gpmcmc.py:193-208, 217-220,
253-264, 288) and PyMC's conventions around it:
  * variable creation order  gv, l, kv, iwgp, cwgp_pos, cwgp, alpha  (= order of the unconstrained vector);
  * HalfNormal / LogNormal get the log transform (``*_log__``), Truncated* the interval transform
    (``*_interval__``), Normal none;
  * ``find_MAP`` maximises logp WITHOUT the transform Jacobian, NUTS samples logp WITH it;
  * initial point = distribution moment (LogNormal exp(mu+sigma^2/2), HalfNormal sigma, Normal mu,
    doubly-truncated (lower+upper)/2).
Everything here is host-side NumPy over a leading batch axis: O(B*P) work next to O(B*N^3) on the device.
"""
import numpy as np

SQRT2PI_LOG = 0.5 * np.log(2.0 * np.pi)


def _norm_cdf(x):
    from scipy.special import ndtr
    return ndtr(x)


class _Block:
    """one named random variable (vector valued)."""

    def __init__(self, name, size, family, args, theta_index):
        self.name, self.size, self.family, self.args = name, size, family, args
        self.theta_index = np.asarray(theta_index, dtype=np.int64)   # where its entries live in the flat theta
        if family in ('lognormal', 'halfnormal'):
            self.transform, self.tname = 'log', name + '_log__'
        elif family in ('truncnormal', 'uniform'):
            self.transform, self.tname = 'interval', name + '_interval__'
        else:
            self.transform, self.tname = None, name

    # -- constrained <-> unconstrained ------------------------------------------------------
    def forward(self, x):
        if self.transform == 'log':
            return np.log(x)
        if self.transform == 'interval':
            lo, hi = self.args[2], self.args[3]
            return np.log(x - lo) - np.log(hi - x)
        return x

    def backward(self, z):
        """x(z), dx/dz, log|dx/dz|, d log|dx/dz| / dz"""
        with np.errstate(all='ignore'):      # far-out proposals of the samplers overflow harmlessly to inf/0
            return self._backward(z)

    def _backward(self, z):
        if self.transform == 'log':
            x = np.exp(z)
            return x, x, z, np.ones_like(z)
        if self.transform == 'interval':
            lo, hi = self.args[2], self.args[3]
            s = 1.0 / (1.0 + np.exp(-z))
            x = lo + (hi - lo) * s
            ljac = np.log(hi - lo) + np.log(s) + np.log1p(-s)
            return x, (hi - lo) * s * (1.0 - s), ljac, 1.0 - 2.0 * s
        return z, np.ones_like(z), np.zeros_like(z), np.zeros_like(z)

    # -- prior log density in the constrained space and its derivative ----------------------
    def logp(self, x):
        with np.errstate(all='ignore'):
            return self._logp(x)

    def _logp(self, x):
        f, a = self.family, self.args
        if f == 'lognormal':
            mu, sg = a
            lx = np.log(x)
            return -0.5 * ((lx - mu) / sg) ** 2 - np.log(sg) - SQRT2PI_LOG - lx, -(lx - mu) / (sg * sg * x) - 1.0 / x
        if f == 'halfnormal':
            sg = a[0]
            return -0.5 * (x / sg) ** 2 + 0.5 * np.log(2.0 / np.pi) - np.log(sg), -x / (sg * sg)
        if f == 'normal':
            mu, sg = a
            return -0.5 * ((x - mu) / sg) ** 2 - np.log(sg) - SQRT2PI_LOG, -(x - mu) / (sg * sg)
        if f == 'truncnormal':
            mu, sg, lo, hi = a
            norm = np.log(_norm_cdf((hi - mu) / sg) - _norm_cdf((lo - mu) / sg))
            return -0.5 * ((x - mu) / sg) ** 2 - np.log(sg) - SQRT2PI_LOG - norm, -(x - mu) / (sg * sg)
        if f == 'uniform':      # args (unused, unused, lower, upper): same slots as truncnormal for the transform
            return np.full_like(x, -np.log(a[3] - a[2])), np.zeros_like(x)
        raise ValueError(f)

    def moment(self):
        f, a = self.family, self.args
        if f == 'lognormal':
            return np.full(self.size, np.exp(a[0] + 0.5 * a[1] ** 2))
        if f == 'halfnormal':
            return np.full(self.size, a[0])
        if f == 'normal':
            return np.full(self.size, a[0])
        return np.full(self.size, 0.5 * (a[2] + a[3]))

    def draw(self, rng, shape):
        f, a = self.family, self.args
        if f == 'lognormal':
            return np.exp(a[0] + a[1] * rng.standard_normal(shape))
        if f == 'halfnormal':
            return np.abs(a[0] * rng.standard_normal(shape))
        if f == 'normal':
            return a[0] + a[1] * rng.standard_normal(shape)
        if f == 'uniform':
            return a[2] + (a[3] - a[2]) * rng.uniform(size=shape)
        from scipy.stats import truncnorm
        lo, hi = (a[2] - a[0]) / a[1], (a[3] - a[0]) / a[1]
        return truncnorm(lo, hi, loc=a[0], scale=a[1]).rvs(size=shape, random_state=rng)


class ParamSpace:
    """All hyperparameters of one model.  ``pos``: positivity flags of the output-warp parameters in wgp order
    (``wgp.pos``), or None when the output warp is not learnable."""
    VECTOR_NAMES = ('l', 'kv', 'iwgp', 'cwgp_pos', 'cwgp')   # reported as arrays even when they hold one entry

    def __init__(self, nx, nkern, noise, n_iw=0, cw_pos=None, has_alpha=False, truncate=False):
        self.nx, self.nkern, self.noise, self.truncate = nx, nkern, noise, truncate
        blocks = []
        p = 0
        if noise:
            fam = ('truncnormal', (0.0, 1e-3, 1e-15, 1.0)) if truncate else ('halfnormal', (1e-3,))
            blocks.append(_Block('gv', 1, fam[0], fam[1], [p]))
            p += 1
        nl = nx * nkern
        fam = ('truncnormal', (0.5, 0.15, 1e-3, 100.0)) if truncate else ('lognormal', (0.0, 1.0))
        blocks.append(_Block('l', nl, fam[0], fam[1], np.arange(p, p + nl)))
        p += nl
        fam = ('truncnormal', (1.0, 0.15, 1e-1, 100.0)) if truncate else ('lognormal', (0.56, 0.75))
        blocks.append(_Block('kv', nkern, fam[0], fam[1], np.arange(p, p + nkern)))
        p += nkern
        if n_iw:
            fam = ('truncnormal', (1.0, 1.0, 1e-3, 5.0)) if truncate else ('lognormal', (0.0, 0.25))
            blocks.append(_Block('iwgp', n_iw, fam[0], fam[1], np.arange(p, p + n_iw)))
            p += n_iw
        self.cw_pos = None
        if cw_pos is not None and len(cw_pos):
            cw_pos = np.asarray(cw_pos, dtype=bool)
            self.cw_pos = cw_pos
            ip, ifree = p + np.where(cw_pos)[0], p + np.where(~cw_pos)[0]
            if len(ip):
                fam = ('truncnormal', (1.0, 1.0, 1e-3, 5.0)) if truncate else ('lognormal', (0.0, 0.25))
                blocks.append(_Block('cwgp_pos', len(ip), fam[0], fam[1], ip))
            if len(ifree):
                fam = ('truncnormal', (0.0, 1.0, -10.0, 10.0)) if truncate else ('normal', (0.0, 1.0))
                blocks.append(_Block('cwgp', len(ifree), fam[0], fam[1], ifree))
            p += len(cw_pos)
        if has_alpha:
            blocks.append(_Block('alpha', 1, 'lognormal', (0.56, 0.75), [p]))
            p += 1
        self._set_blocks(blocks, p)

    def _set_blocks(self, blocks, p):
        self.blocks = blocks
        self.P = p
        # unconstrained vector: blocks concatenated in creation order
        off, self.zslices = 0, []
        for b in blocks:
            self.zslices.append(slice(off, off + b.size))
            off += b.size
        assert off == p

    # ---- conversions -----------------------------------------------------------------------
    def initial_z(self, start=None):
        """PyMC initial point (moments), optionally overridden by a hypers-style dict (constrained or
        transformed names both accepted, as ``find_MAP(start=...)`` does)."""
        z = np.empty(self.P)
        for b, sl in zip(self.blocks, self.zslices):
            x0 = b.moment()
            if start is not None:
                if b.tname in start and b.transform is not None:
                    z[sl] = np.asarray(start[b.tname], dtype=np.float64).reshape(-1)
                    continue
                if b.name in start:
                    x0 = np.asarray(start[b.name], dtype=np.float64).reshape(-1)
            z[sl] = b.forward(x0)
        return z

    def theta_from_z(self, z):
        """z [.., P] -> theta [.., P] in the flat device layout, plus dtheta/dz, log-Jacobian and its z-gradient."""
        z = np.asarray(z, dtype=np.float64)
        theta = np.empty_like(z)
        dxdz = np.empty_like(z)
        ljac = np.zeros(z.shape[:-1])
        dljac = np.zeros_like(z)
        for b, sl in zip(self.blocks, self.zslices):
            x, dx, lj, dlj = b.backward(z[..., sl])
            theta[..., b.theta_index] = x
            dxdz[..., sl] = dx
            ljac = ljac + lj.sum(axis=-1)
            dljac[..., sl] = dlj
        return theta, dxdz, ljac, dljac

    def theta_only(self, z):
        """z [.., P] -> theta [.., P]: the first output of :meth:`theta_from_z` alone (same values), for callers that launch
        the device evaluation before they form the Jacobian terms."""
        z = np.asarray(z, dtype=np.float64)
        theta = np.empty_like(z)
        with np.errstate(all='ignore'):
            for b, sl in zip(self.blocks, self.zslices):
                zb = z[..., sl]
                if b.transform == 'log':
                    theta[..., b.theta_index] = np.exp(zb)
                elif b.transform == 'interval':
                    lo, hi = b.args[2], b.args[3]
                    theta[..., b.theta_index] = lo + (hi - lo) * (1.0 / (1.0 + np.exp(-zb)))
                else:
                    theta[..., b.theta_index] = zb
        return theta

    def z_from_theta(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        z = np.empty_like(theta)
        for b, sl in zip(self.blocks, self.zslices):
            z[..., sl] = b.forward(theta[..., b.theta_index])
        return z

    def prior(self, theta):
        """sum of prior log densities at constrained theta [.., P] and the gradient w.r.t. theta."""
        theta = np.asarray(theta, dtype=np.float64)
        lp = np.zeros(theta.shape[:-1])
        g = np.zeros_like(theta)
        for b in self.blocks:
            v, dv = b.logp(theta[..., b.theta_index])
            lp = lp + v.sum(axis=-1)
            g[..., b.theta_index] = dv
        return lp, g

    def grad_theta_to_z(self, gtheta, dxdz):
        """chain rule: gradient w.r.t. theta (device layout) -> gradient w.r.t. z (block order)."""
        gz = np.empty_like(gtheta)
        for b, sl in zip(self.blocks, self.zslices):
            gz[..., sl] = gtheta[..., b.theta_index] * dxdz[..., sl]
        return gz

    def draw_prior_z(self, rng, n):
        z = np.empty((n, self.P))
        for b, sl in zip(self.blocks, self.zslices):
            z[:, sl] = b.forward(b.draw(rng, (n, b.size)))
        return z

    def hypers_dict(self, z):
        """dict keyed like PyMC's find_MAP result: transformed and constrained names (tutorial.ipynb:529)."""
        z = np.asarray(z, dtype=np.float64).reshape(-1)
        out = {}
        for b, sl in zip(self.blocks, self.zslices):
            if b.transform is not None:
                out[b.tname] = z[sl].copy() if b.size > 1 or b.name in self.VECTOR_NAMES \
                    else np.array(z[sl][0])
        for b, sl in zip(self.blocks, self.zslices):
            x = b.backward(z[sl])[0]
            out[b.name] = x.copy() if b.size > 1 or b.name in self.VECTOR_NAMES else np.array(x[0])
        return out

    def theta_from_hypers(self, hyp):
        th = np.empty(self.P)
        for b in self.blocks:
            th[b.theta_index] = np.asarray(hyp[b.name], dtype=np.float64).reshape(-1)
        return th


class XSpace(ParamSpace):
    """The unknown input point of the inverse problem / the query point of the BO ``opt_method='map'`` graph as
    PyMC variables ``x0 .. x{nx-1}``: the scipy priors of the sampler converted as the reference does
    (andvaranaut/gpmcmc.py:1053-1096 and :706-731): uniform -> pm.Uniform (interval transform), norm -> pm.Normal,
    truncnorm -> pm.TruncatedNormal (interval transform)."""
    VECTOR_NAMES = ()

    def __init__(self, priors):
        blocks = []
        for k, pr in enumerate(priors):
            kind = type(pr.dist).__name__
            if kind == 'uniform_gen':
                lo, hi = pr.support()
                blocks.append(_Block(f'x{k}', 1, 'uniform', (0.0, 1.0, float(lo), float(hi)), [k]))
            elif kind == 'norm_gen':
                blocks.append(_Block(f'x{k}', 1, 'normal', (float(pr.mean()), float(pr.std())), [k]))
            elif kind == 'truncnorm_gen':
                lo, hi = pr.support()
                # the reference recovers mu / sigma from the frozen distribution's args / kwds case by case
                # (gpmcmc.py:1063-1091); scipy's own argument parser gives the same (loc, scale)
                _, loc, scale = pr.dist._parse_args(*pr.args, **pr.kwds)
                blocks.append(_Block(f'x{k}', 1, 'truncnormal', (float(loc), float(scale), float(lo), float(hi)), [k]))
            else:
                raise Exception('Prior distribution conversion from scipy to pymc not implemented')
        self.nx = len(priors)
        self._set_blocks(blocks, len(priors))

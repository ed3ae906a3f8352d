"""GP surrogate with the public surface of andvaranaut's ``GPMCMC`` (andvaranaut/gpmcmc.py:30), driven by the
B200-native engine instead of PyMC/PyTensor.

Kept from the reference (same names, arguments, defaults, error messages where they are part of the contract):
constructor (:31-40), ``set_data`` (:122-137), ``sample`` (:158-172), ``fit`` (:175-182, model semantics of
``__fit`` :185-401), ``predict`` (:522-542), ``BO`` (:601-906), ``inverse_opt`` (:1040-1217), ``change_model``
(:472-519), ``change_conrevs/xconrevs/yconrevs`` (:75-95), ``cwgp_set/iwgp_set`` (:433-462), ``train_test``
(:465-469), ``mean_extract/map_extract`` (:404-430), ``del_samples`` (:57-72), ``relative_importances``,
``test_metrics`` (the numeric half of ``test_plots`` :933-1028; plotting is out of scope).

Replaced: the PyMC model (``self.m``) is a :class:`~andvaranaut_b200.priors.ParamSpace` + :class:`GPEngine`
(``self.gp``); ``pm.find_MAP`` / ``pm.sample`` are :mod:`andvaranaut_b200.drivers` (L-BFGS-B, lock-step NUTS);
``gp.predict`` and the per-point Python loop ``__gh_stats`` (:545-569) are one device call; the PyMC models over the
query point x (BO refine / ``opt_method='map'`` :699-858, ``inverse_opt`` :1049-1165) are
:class:`~andvaranaut_b200.xpost.XPosterior` objects whose potential is evaluated on the device with its analytic
gradient.  Documented deviations: ``seed`` is honoured by ``sample``; ``restarts > 1`` really restarts from different
(prior-drawn) points (the reference builds a start and never passes it on, :330-332); the factorisation behind
``predict`` is cached and extended by rank-1 appends instead of being rebuilt on every call (:588-598).
"""
import copy
import re
from time import time as stopwatch

import numpy as np

from . import transform as T
from .drivers import Posterior, find_map, find_map_multi, sample as hmc_sample
from .lhc import LHC
from .priors import ParamSpace
from .xpost import InverseLikelihood, XPosterior, kdiag_values
from .transform import wgp

__all__ = ['GPMCMC']

KERNELS = ['RBF', 'Matern52', 'Matern32', 'Exponential', 'RatQuad']


class _none_conrev:
    """identity conversion (gpmcmc.py:23-27)"""

    def con(self, x):
        return x

    def rev(self, x):
        return x


def _frozen_programs(obj):
    """(forward program, inverse program) of a conrev object if the device can evaluate it, else None."""
    if isinstance(obj, _none_conrev):
        return [], []
    if isinstance(obj, wgp):
        p = obj.rev_program()
        return p, p
    st = T.frozen_stage(obj)
    if st is not None:
        return [st], [st]
    return None


class GPMCMC(LHC):
    def __init__(self, xconrevs=None, yconrevs=None, kernel='RBF', noise=True, mean=0, device=None, **kwargs):
        super().__init__(**kwargs)
        self.device = device
        self.xc = copy.deepcopy(self.x)
        self.yc = copy.deepcopy(self.y)
        self.__conrev_check(xconrevs, yconrevs)
        self.change_model(kernel, noise, mean)
        self.__scrub_train_test()
        self.ym = copy.deepcopy(self.y)
        self.xopt = self.yopt = None
        self.shard = None          # optional andvaranaut_b200.dist.Shard for multi-GPU runs

    # ---- pickling (save_object / load_object, core.py): device state is dropped and rebuilt on the next predict ----
    def __getstate__(self):
        st = dict(self.__dict__)
        st['gp'] = None
        st['_pred_cache'] = None
        st['shard'] = None
        return st

    # ---- small pieces of state management --------------------------------------------------------
    def zero_mean(self, x):
        return np.zeros(self.ny)

    def __mean_values(self, x):
        if self.mean == self.zero_mean:
            return np.zeros((len(x), self.ny))
        xm, ym = self._core__vector_solver(x, self.mean)
        if len(xm) != len(x):
            raise Exception('Mean function not valid at every x point in dataset')
        return ym

    def __con(self, nsamps):
        self.xc = np.r_[self.xc, np.zeros((nsamps, self.nx))]
        self.yc = np.r_[self.yc, np.zeros((nsamps, self.ny))]
        for i in range(self.nx):
            self.xc[-nsamps:, i] = self.xconrevs[i].con(self.x[-nsamps:, i])
        for i in range(self.ny):
            self.yc[-nsamps:, i] = self.yconrevs[i].con(self.y[-nsamps:, i] - self.ym[-nsamps:, i])

    def del_samples(self, ndels=None, method='coarse_lhc', idx=None):
        returned = super()._LHC__del_samples(ndels, method, idx, returns=True)
        if method == 'coarse_lhc':
            for i in range(ndels):
                self.xc = np.delete(self.xc, returned[i], axis=0)
                self.yc = np.delete(self.yc, returned[i], axis=0)
                self.ym = np.delete(self.ym, returned[i], axis=0)
        else:
            self.xc, self.yc, self.ym = self.xc[returned], self.yc[returned], self.ym[returned]
        self.nsamp = len(self.x)
        self.__scrub_train_test()
        self._pred_cache = None

    def change_conrevs(self, xconrevs=None, yconrevs=None):
        self.__conrev_check(xconrevs, yconrevs)
        for i in range(self.nx):
            self.xc[:, i] = self.xconrevs[i].con(self.x[:, i])
        for i in range(self.ny):
            self.yc[:, i] = self.yconrevs[i].con(self.y[:, i] - self.ym[:, i])
        self._pred_cache = None

    def change_xconrevs(self, xconrevs=None):
        self.__conrev_check(xconrevs, yconrevs=self.yconrevs)
        for i in range(self.nx):
            self.xc[:, i] = self.xconrevs[i].con(self.x[:, i])
        self._pred_cache = None

    def change_yconrevs(self, yconrevs=None):
        self.__conrev_check(self.xconrevs, yconrevs)
        for i in range(self.ny):
            self.yc[:, i] = self.yconrevs[i].con(self.y[:, i] - self.ym[:, i])
        self._pred_cache = None

    def __conrev_check(self, xconrevs, yconrevs):
        xconrevs = [None] * self.nx if xconrevs is None else xconrevs
        yconrevs = [None] * self.ny if yconrevs is None else yconrevs
        if not isinstance(xconrevs, list) or len(xconrevs) != self.nx:
            raise Exception('Error: xconrevs must be None or list of conversion/reversion classes of size nx')
        if not isinstance(yconrevs, list) or len(yconrevs) != self.ny:
            raise Exception('Error: yconrevs must be None or list of conversion/reversion classes of size ny')
        xconrevs, yconrevs = list(xconrevs), list(yconrevs)
        for lst in (xconrevs, yconrevs):
            for j, c in enumerate(lst):
                if c is None:
                    lst[j] = _none_conrev()
                elif not callable(c.con) or not callable(c.rev):
                    raise Exception('Error: Provided data conversion/reversion function not callable.')
        self.xconrevs, self.yconrevs = xconrevs, yconrevs

    def set_data(self, x, y):
        super().set_data(x, y)
        self.xc = np.empty((0, self.nx))
        self.yc = np.empty((0, self.ny))
        self.ym = self.__mean_values(self.x)
        self.__con(self.nsamp)
        self.__scrub_train_test()
        self._pred_cache = None

    def __scrub_train_test(self):
        self.train = None
        self.test = None

    def sample(self, nsamps, seed=None):
        super().sample(nsamps=nsamps, seed=seed)
        self.ym = self.__mean_values(self.x)
        self.xc = np.empty((0, self.nx))
        self.yc = np.empty((0, self.ny))
        self.nsamp = len(self.x)
        self.__con(self.nsamp)
        self._pred_cache = None

    def train_test(self, training_frac=0.9, seed=None):
        self.nsamp = len(self.x)
        rng = np.random.default_rng(seed)
        perm = rng.permutation(self.nsamp)
        ntr = int(round(training_frac * self.nsamp))
        self.train, self.test = perm[:ntr], perm[ntr:]

    def change_model(self, kernel=None, noise=None, mean=None):
        kernel = self.kernel if kernel is None else kernel
        noise = self.noise if noise is None else noise
        if mean is not None:
            self.mean = self.zero_mean if (not callable(mean) and mean == 0) else mean
            self.ym = self.__mean_values(self.x)
        kerns = re.split(r'[+*]', kernel)
        ops = [c for c in kernel if c in '+*']
        for k in kerns:
            if k not in KERNELS:
                raise Exception(f'Error: kernel string must contain only {KERNELS}')
        if kerns.count('RatQuad') > 1:
            raise Exception('Error: only one RatQuad kernel may be specified')
        if not isinstance(noise, bool):
            raise Exception('Error: noise must be of type bool')
        self.kernel, self.kerns, self.ops, self.nkern, self.noise = kernel, kerns, ops, len(kerns), noise
        self.m = None
        self.gp = None
        self.hypers = None
        self._pred_cache = None

    # ---- warps -------------------------------------------------------------------------------------
    def cwgp_set(self, params, mode='numpy', y=None):
        y = self.y - self.ym if y is None else y
        warper = wgp(self.yconrevs[0].warping_names, params, y[:, 0])
        if mode == 'numpy':
            self.change_yconrevs([warper])
        else:
            return warper

    def iwgp_set(self, params, mode='numpy', x=None):
        x = self.x if x is None else x
        out, rc = [], 0
        for i in range(self.nx):
            c = self.xconrevs[i]
            if isinstance(c, wgp):
                n = len(c.params)
                out.append(wgp(c.warping_names, params[rc:rc + n], y=x[:, i], xdist=self.priors[i]))
                rc += n
            else:
                out.append(c)
        if mode == 'numpy':
            self.change_xconrevs(xconrevs=out)
        else:
            return out

    # ---- fit ---------------------------------------------------------------------------------------
    def fit(self, method='map', return_data=False, iwgp=False, cwgp=False, jitter=1e-6, truncate=False,
            restarts=1, **kwargs):
        self.m, self.gp, self.hypers, data = self.__fit(self.x, self.y - self.ym, method, iwgp, cwgp, jitter,
                                                        truncate, restarts, **kwargs)
        if method != 'none':       # unchanged hypers and conrevs: the factorised state stays valid (and is EXTENDED by
            self._pred_cache = None  # rank-1 appends when points were added since, see _predict_engine)
        if return_data:
            return data

    def _build_model(self, x, y, iwgp, cwgp, jitter, truncate):
        """engine + parameter space for (x raw, y raw minus mean): the model ``__fit`` declares (:189-323)."""
        from .gp import GPEngine
        n_iw, xprogs = 0, None
        xin = np.empty_like(x)
        if iwgp:
            xprogs = []
            for i in range(self.nx):
                c = self.xconrevs[i]
                if isinstance(c, wgp):
                    xprogs.append(c.program())
                    xin[:, i] = x[:, i]                      # raw: warped on the device per hyperparameter sample
                    n_iw += c.np
                else:
                    xprogs.append(None)
                    xin[:, i] = c.con(x[:, i])
            if n_iw == 0:
                raise Exception('Error: iwgp set to true but none of xconrevs are wgp classes')
        else:
            for i in range(self.nx):
                xin[:, i] = self.xconrevs[i].con(x[:, i])
        yprog, cw_pos = None, None
        if cwgp:
            c = self.yconrevs[0]
            if not isinstance(c, wgp):
                raise Exception('Error: cwgp set to true but yconrevs class is not wgp')
            if c.np == 0:
                raise Exception('Error: cwgp set to true but wgp class has no tuneable parameters')
            yprog, cw_pos = c.program(), c.pos
            yin = y[:, 0]
        else:
            yin = self.yconrevs[0].con(y[:, 0])
        eng = GPEngine(nx=self.nx, kerns=self.kerns, ops=self.ops, noise=self.noise, jitter=jitter, xwarps=xprogs,
                       ywarp=yprog, device=self.device)
        eng.set_data(xin, yin)
        space = ParamSpace(self.nx, self.nkern, self.noise, n_iw=n_iw, cw_pos=cw_pos,
                           has_alpha='RatQuad' in self.kerns, truncate=truncate)
        assert space.P == eng.P
        return eng, space

    def __fit(self, x, y, method, iwgp, cwgp, jitter=1e-6, truncate=False, restarts=1, **kwargs):
        if len(y) < 1:
            raise Exception('Error: no data to fit; call sample() or set_data() first')
        eng, space = self._build_model(x, y, iwgp, cwgp, jitter, truncate)
        post = Posterior(eng, space, shard=self.shard)
        start = kwargs.pop('start', None)
        maxeval = kwargs.pop('maxeval', 5000)
        kwargs.pop('progressbar', None)
        data = None
        if method == 'map':
            z0 = space.initial_z(start)
            if restarts > 1:
                rng = np.random.default_rng(kwargs.pop('seed', None))
                z0s = np.vstack([z0[None, :], space.draw_prior_z(rng, restarts - 1)])
                zs, lps = find_map_multi(post, z0s, maxeval=maxeval, **kwargs)
                zbest = zs[int(np.nanargmax(np.where(np.isfinite(lps), lps, -np.inf)))]
                data = {'z': zs, 'logp': lps}
            else:
                kwargs.pop('seed', None)
                zbest, lp, nev = find_map(post, z0, maxeval=maxeval, **kwargs)
                data = {'logp': lp, 'evals': nev}
                if self.verbose:
                    print(f'MAP: logp = {lp:,.5g}, {nev} evaluations')
            mp = space.hypers_dict(zbest)
        elif method == 'none':
            mp = self.hypers
            if mp is None:
                raise Exception("Error: method='none' needs previously fitted hypers")
        elif method in ('mcmc_mean', 'mcmc_map'):
            skw = {k: kwargs.pop(k) for k in ('draws', 'tune', 'chains', 'seed', 'target_accept', 'max_leapfrog',
                                              'path_length', 'init_jitter', 'sampler', 'max_treedepth') if k in kwargs}
            skw.setdefault('chains', 4)
            kwargs.pop('cores', None)
            kwargs.pop('random_seed', None)
            data = hmc_sample(post, start_z=None if start is None else space.initial_z(start)[None, :], **skw)
            if method == 'mcmc_mean':
                mp = self.mean_extract(data)
                # the posterior mean of each constrained variable defines its transformed twin
                z = space.z_from_theta(space.theta_from_hypers(mp))
                mp = space.hypers_dict(z)
            else:
                mp = self.map_extract(data)
                try:       # the reference's polish may fail numerically (gpmcmc.py:356-361); device errors propagate
                    z, _, _ = find_map(post, space.initial_z(mp), maxeval=maxeval)
                    mp = space.hypers_dict(z)
                except (ValueError, FloatingPointError, np.linalg.LinAlgError):
                    pass
        else:
            raise Exception('method must be one of map, mcmc_map, or mcmc_mean')

        # bake learnt warps into the NumPy conrevs and refresh the converted caches (gpmcmc.py:364-399)
        if method != 'none':
            if iwgp:
                self.iwgp_set(np.asarray(mp['iwgp']).reshape(-1))
            if cwgp:
                params, ip, ifr = [], 0, 0
                for flag in self.yconrevs[0].pos:
                    if flag:
                        params.append(np.asarray(mp['cwgp_pos']).reshape(-1)[ip])
                        ip += 1
                    else:
                        params.append(np.asarray(mp['cwgp']).reshape(-1)[ifr])
                        ifr += 1
                self.cwgp_set(np.array(params))
        return space, None, mp, data

    def mean_extract(self, data):
        mp = {}
        for key, v in data.posterior.items():
            mp[key] = np.asarray(v).mean(axis=(0, 1))
        return mp

    def map_extract(self, data):
        lp = np.asarray(data.sample_stats['lp'])
        c, d = np.unravel_index(int(np.argmax(lp)), lp.shape)
        if self.verbose:
            print(f'Max log posterior: {lp[c, d]}')
        return {key: np.asarray(v)[c, d] for key, v in data.posterior.items()}

    # ---- predict -----------------------------------------------------------------------------------
    def _predict_engine(self, jitter):
        """engine over the CONVERTED data (xc, yc) with the fitted hypers factorised once; cached until data,
        conrevs, model or hypers change (the reference refactorises and recompiles on every call, :588-598)."""
        if self.hypers is None:
            raise Exception('Error: model must be fitted before predicting')
        from .gp import GPEngine, check_info
        space = ParamSpace(self.nx, self.nkern, self.noise, has_alpha='RatQuad' in self.kerns)
        th = space.theta_from_hypers(self.hypers)
        n = len(self.xc)
        c = self._pred_cache
        # the cache is valid for exactly the state it was built from: jitter, the hyperparameter VALUES (the reference
        # passes point=self.hypers on every call, so edited / loaded hypers must take effect) and the converted data
        # (a checksum of xc / yc catches in-place edits at unchanged length)
        if c is not None and c['jitter'] == jitter and np.array_equal(c['theta'], th):
            n0 = c['n']
            if n0 <= n and n - n0 <= 64 and c['sum'] == self.__data_sum(n0):
                if n == n0:
                    return c['eng'], th
                # points were appended (BO / inverse_opt / fit_method='none') under unchanged hypers and conversions:
                # rank-1 extension of the cached factorisation, O(N^2) per point
                eng, ok = c['eng'], True
                for i in range(n0, n):
                    info = int(eng.append(self.xc[i], self.yc[i, 0])[0])
                    check_info(info, 'avn_gp_append')
                    if info != 0:
                        ok = False
                        break
                if ok:
                    c.update(n=n, sum=self.__data_sum(n))
                    return eng, th
        self._pred_cache = None
        eng = GPEngine(nx=self.nx, kerns=self.kerns, ops=self.ops, noise=self.noise, jitter=jitter, device=self.device)
        eng.set_data(self.xc, self.yc[:, 0])
        info = int(eng.factorize(th)[0])
        check_info(info, 'avn_gp_factorize')
        if info != 0:
            raise Exception(f'Error: covariance matrix not positive definite at pivot {info}')
        self.gp = eng
        self._pred_cache = dict(jitter=jitter, theta=th.copy(), n=n, sum=self.__data_sum(n), eng=eng)
        return eng, th

    def __data_sum(self, n):
        """order-dependent checksum of the first n converted training rows (bit patterns, so -0.0 / NaN edits count)."""
        import zlib
        return (zlib.crc32(np.ascontiguousarray(self.xc[:n]).view(np.uint8)),
                zlib.crc32(np.ascontiguousarray(self.yc[:n, 0]).view(np.uint8)))

    def predict(self, x, return_var=False, convert=True, revert=True, normvar=False, jitter=1e-6, EI=False,
                EIopt=None, deg=8):
        from .gp import GPEngine
        x = np.asarray(x, dtype=np.float64)
        if convert:
            xarg = np.zeros_like(x)
            for i in range(self.nx):
                xarg[:, i] = self.xconrevs[i].con(x[:, i])
        else:
            xarg = copy.deepcopy(x)
            for i in range(self.nx):
                x[:, i] = self.xconrevs[i].rev(x[:, i])      # reference semantics: caller's array becomes raw x
        t0 = stopwatch()
        eng, _ = self._predict_engine(jitter)
        progs = _frozen_programs(self.yconrevs[0]) if revert else None
        madd = None
        if revert and self.mean != self.zero_mean:
            madd = self.__mean_values(x)[:, 0]
        if revert and progs is not None:
            epi = GPEngine.make_epilogue(mode='EI' if EI else 'revert', deg=deg, normvar=normvar, EIopt=EIopt,
                                         yopt=0.0 if self.yopt is None else float(self.yopt), yrev=progs[1])
            mu, var = self._dev_predict(eng, xarg, epi, madd)
        else:
            mu, var = self._dev_predict(eng, xarg, GPEngine.make_epilogue(mode='latent'), None)
            if revert:   # user-defined y transform: reversion on the host, vectorised over points
                mu, var = self.__gh_stats_host(mu, var, madd, normvar, deg, EI, EIopt)
        if self.verbose:
            print(f'Predicting...\nTime taken: {stopwatch() - t0:0.2f} s')
        y, yv = mu.reshape(-1, 1), var.reshape(-1, 1)
        return (y, yv) if return_var else y

    def _dev_predict(self, eng, xarg, epi, madd):
        if self.shard is not None and self.shard.world > 1:
            return self.shard.predict(eng, xarg, epilogue=epi, mean_add=madd)
        mu, var = eng.predict(xarg, epilogue=epi, mean_add=madd)
        return mu.cpu().numpy(), var.cpu().numpy()

    def __gh_stats_host(self, mu, var, madd, normvar, deg, EI, EIopt):
        xi, wi = np.polynomial.hermite.hermgauss(deg)
        yi = np.sqrt(2 * var)[:, None] * xi[None, :] + mu[:, None]
        yir = np.stack([self.yconrevs[0].rev(yi[:, k]) for k in range(deg)], axis=1)
        if madd is not None:
            yir = yir + madd[:, None]
        f = yir
        if EI:
            dff = yir - self.yopt if EIopt == 'max' else self.yopt - yir
            f = np.where(dff > 0.0, dff, 0.0)
        m = np.sum(wi * f, axis=1) / np.sqrt(np.pi)
        v = np.sum(wi * yir ** 2, axis=1) / np.sqrt(np.pi) - m ** 2
        if normvar:
            v = v / m ** 2
        return m, v

    # ---- diagnostics -------------------------------------------------------------------------------
    def test_metrics(self, revert=True, iwgp=False, cwgp=False, method='none'):
        """RMSE / MAE / mean fractional error / R^2 on the held-out split: the numbers ``test_plots`` prints
        (gpmcmc.py:933-976): fit on the training part (``method='none'`` reuses the current hypers), predict the
        test part, compare in original (revert) or converted units."""
        if self.train is None:
            self.train_test()
        full = (self.x, self.y, self.ym, self.xc, self.yc)
        saved = (self.m, self.hypers)
        tr, te = self.train, self.test
        try:
            self.x, self.y, self.ym, self.xc, self.yc = (a[tr] for a in full)
            self._pred_cache = None
            if method != 'none':
                self.fit(method=method, iwgp=iwgp, cwgp=cwgp)
            pred, pvar = self.predict(full[0][te].copy(), return_var=True, revert=revert)
        finally:
            self.x, self.y, self.ym, self.xc, self.yc = full
            self.m, self.hypers = saved
            self._pred_cache = None
        if revert:
            truth, meany = full[1][te][:, 0], np.mean(full[1])
        else:
            truth, meany = self.yconrevs[0].con(full[1][te][:, 0] - full[2][te][:, 0]), np.mean(full[4])
        err = pred[:, 0] - truth
        out = dict(rmse=float(np.sqrt(np.mean(err ** 2))), mae=float(np.mean(np.abs(err))),
                   mpe=float(np.mean(np.abs(err) / np.abs(truth))),
                   r2=float(1 - np.sum(err ** 2) / np.sum((truth - meany) ** 2)))
        if self.verbose:
            print(f"RMSE for y is: {out['rmse']:0.5e}")
            print(f"Mean absoulte error for y is: {out['mae']:0.5e}")
            print(f"Mean percentage error for y is: {out['mpe']:0.5%}")
            print(f"R^2 for y is: {out['r2']:0.5f}")
        return out

    def y_dist(self, mode='hist_kde', nsamps=None, return_data=False, surrogate=True, seed=None):
        """distribution of the output over the input priors (gpmcmc.py:140-151): ``nsamps`` LHC points through the
        surrogate (one batched device predict) or the stored data; the seaborn plot of the reference is drawn only when
        seaborn / matplotlib are installed (plotting is out of scope here), the data come back with ``return_data``."""
        if mode not in ('hist', 'kde', 'ecdf', 'hist_kde'):
            raise Exception("Error: selected mode must be one of ['hist', 'kde', 'ecdf', 'hist_kde']")
        if not isinstance(surrogate, bool):
            raise Exception('Error: surrogate argument must be of type bool')
        if surrogate:
            xs = self._LHC__latin_sample(nsamps, seed=seed)
            ys = self.predict(xs)
        else:
            xs, ys = self.x, self.y
        try:
            import seaborn as sns
            import matplotlib.pyplot as plt
            kw = dict(hist=dict(kind='hist'), kde=dict(kind='kde'), ecdf=dict(kind='ecdf'), hist_kde=dict(kind='hist', kde=True))
            for i in range(self.ny):
                sns.displot(ys[:, i], **kw[mode])
                plt.xlabel(f'y[{i}]')
                plt.ylabel('Density')
                plt.show()
        except ImportError:
            pass
        if return_data and surrogate:
            return xs, ys

    def relative_importances(self):
        ls = np.asarray(self.hypers['l']).reshape(self.nkern, self.nx)
        inv = 1.0 / ls
        return inv / inv.sum(axis=1, keepdims=True)

    # ---- Bayesian optimisation ---------------------------------------------------------------------
    def BO(self, opt_type='min', opt_method='predict', fit_method='map', max_iter=16, method='EI', eps=0.1,
           iwgp=False, cwgp=False, jitter=1e-6, conv=0.01, predict_samps=10000, normvar=True, refine=True,
           seed=None, **kwargs):
        if self.ny > 1:
            raise Exception('Bayesian minimisation only implemented for single output')
        if opt_type == 'max':
            xoptf, yoptf = np.argmax, np.max
        elif opt_type == 'min':
            xoptf, yoptf = np.argmin, np.min
        else:
            raise Exception('Error: opt_type argument must be one of max or min')
        if method not in ('eps-RS', 'exploit', 'explore', 'EI'):
            raise Exception('method must be one of eps-RS ,EI, exploit, or explore')
        self.xopt = self.x[xoptf(self.y[:, 0]), :]
        self.yopt = yoptf(self.y)
        if self.verbose:
            print('Running Bayesian minimisation...')
            print(f'Current optima is {self.yopt} at x point {self.xopt}')
        if self.hypers is None:
            raise Exception('Model must be fitted before running Bayesian optimisation')
        if method == 'exploit':
            eps = 0.0
        rng = np.random.default_rng(seed)
        lbs = np.array([p.ppf(1e-8) for p in self.priors])
        ubs = np.array([p.isf(1e-8) for p in self.priors])
        verb = self.verbose

        def optf(x):
            x = np.atleast_2d(x)
            self.verbose = False
            try:
                if method in ('eps-RS', 'exploit'):
                    ym = self.predict(x, jitter=jitter)
                    return ym[:, 0] if opt_type == 'min' else -ym[:, 0]
                if method == 'explore':
                    _, yv = self.predict(x, return_var=True, normvar=normvar, jitter=jitter)
                    return -yv[:, 0]
                ym = self.predict(x, EI=True, EIopt=opt_type, jitter=jitter)
                return -ym[:, 0]
            finally:
                self.verbose = verb

        xsampold = np.full((1, self.nx), 1e300)
        for it in range(max_iter):
            if self.verbose:
                print(f'Iteration {it + 1}')
            roll = rng.uniform()
            if method != 'eps-RS' or roll > eps:
                if opt_method == 'DE':
                    from scipy.optimize import differential_evolution
                    res = differential_evolution(optf, list(zip(lbs, ubs)), vectorized=True, updating='deferred',
                                                 seed=int(rng.integers(2 ** 31)))
                    xsamp = np.array([res.x])
                elif opt_method in ('map', 'mcmc_mean', 'mcmc_map'):
                    xsamp = self.__acquisition_model_opt(opt_method, method, opt_type, normvar, jitter, rng, **kwargs)
                elif opt_method != 'predict':
                    raise Exception('opt_method must be one of map, mcmc_map, or mcmc_mean')
                else:
                    xs = self._LHC__latin_sample(predict_samps, seed=int(rng.integers(2 ** 31)))
                    ys = optf(xs)
                    xsamp = np.array([xs[int(np.argmin(ys)), :]])
                    if self.verbose:
                        print(f'Function opt is {np.min(ys):0.3f}')
                if refine and opt_method == 'predict':
                    xsamp = self.__refine(optf, xsamp, lbs, ubs,
                                          acq=lambda xx: self.acquisition_grad(xx, method=method, opt_type=opt_type,
                                                                               normvar=normvar, jitter=jitter))
            else:
                xsamp = np.array([[p.rvs(random_state=rng) for p in self.priors]])
            xdiff = np.sum(np.abs(xsamp - xsampold) / np.abs(xsampold)) / self.nx
            if xdiff < conv:
                if self.verbose:
                    print(f'Convergence at relative tolerance {xdiff} achieved with point {xsamp}')
                break
            xsampold = xsamp
            xs_new, ys_new = self._core__vector_solver(xsamp)
            if len(xs_new) == 0:
                continue
            ym_new = self.__mean_values(xs_new)
            self.x = np.r_[self.x, xs_new]
            self.y = np.r_[self.y, ys_new]
            self.ym = np.r_[self.ym, ym_new]
            self.nsamp = len(self.x)
            self.__con(len(xs_new))
            if self.verbose:
                print(f'New sample is {ys_new} at x point {xs_new}')
            self.xopt = self.x[xoptf(self.y[:, 0]), :]
            self.yopt = yoptf(self.y)
            if fit_method == 'map':
                from .gp import GPError
                try:       # warm start at the previous hypers; the reference retries from the default start (:897-904)
                    self.fit(method=fit_method, iwgp=iwgp, cwgp=cwgp, jitter=jitter, start=self.hypers)
                except GPError:
                    raise
                except Exception:
                    self.fit(method=fit_method, iwgp=iwgp, cwgp=cwgp, jitter=jitter)
            else:
                self.fit(method=fit_method, iwgp=iwgp, cwgp=cwgp, jitter=jitter, **kwargs)
        return self.xopt, self.yopt

    def _con_with_der(self, x):
        """converted inputs and d con / d x per column (analytic ``der`` where the conrev class has one, central
        differences of ``con`` otherwise)."""
        xc, dc = np.empty_like(x), np.empty_like(x)
        for i in range(self.nx):
            c = self.xconrevs[i]
            xc[:, i] = c.con(x[:, i])
            if isinstance(c, _none_conrev):
                dc[:, i] = 1.0
            elif hasattr(c, 'der'):
                dc[:, i] = c.der(x[:, i])
            else:
                h = 1e-6 * np.maximum(1.0, np.abs(x[:, i]))
                dc[:, i] = (c.con(x[:, i] + h) - c.con(x[:, i] - h)) / (2 * h)
        return xc, dc

    def acquisition_grad(self, x, method='EI', opt_type='min', normvar=True, jitter=1e-6, deg=8):
        """Value and gradient w.r.t. the raw query points x [M,nx] of the acquisition the reference builds inline
        for its BO refine / ``opt_method='map'`` step (gpmcmc.py:738-829) and differentiates with PyTensor: the
        reverted predictive mean (exploit / eps-RS), the reverted variance WITHOUT the noise term (explore) or the
        expected improvement (EI), signed so that SMALLER is better.  One device call for all M points.
        Returns (f [M], g [M,nx]), or None when the output transform has no device program."""
        from .gp import GPEngine
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        progs = _frozen_programs(self.yconrevs[0])
        if progs is None:
            return None
        eng, _ = self._predict_engine(jitter)
        xc, dc = self._con_with_der(x)
        madd = dmadd = None
        if self.mean != self.zero_mean:
            madd = self.__mean_values(x)[:, 0]
            dmadd = np.empty_like(x)
            for i in range(self.nx):       # user mean function: central differences on the host
                h = np.zeros(self.nx)
                h[i] = 1e-6 * max(1.0, float(np.max(np.abs(x[:, i]))))
                dmadd[:, i] = (self.__mean_values(x + h)[:, 0] - self.__mean_values(x - h)[:, 0]) / (2 * h[i])
            dmadd = dmadd / dc             # the device differentiates w.r.t. the converted inputs
        epi = GPEngine.make_epilogue(mode='EI' if method == 'EI' else 'revert', deg=deg,
                                     normvar=normvar and method == 'explore', EIopt=opt_type,
                                     yopt=0.0 if self.yopt is None else float(self.yopt), yrev=progs[1])
        m, v, dm, dv = (t.cpu().numpy() for t in eng.predict_grad(xc, epilogue=epi, mean_add=madd, dmean_add=dmadd,
                                                                  pred_noise=False))
        if method in ('eps-RS', 'exploit'):
            sgn = 1.0 if opt_type == 'min' else -1.0
            return sgn * m, sgn * dm * dc
        if method == 'explore':
            return -v, -dv * dc
        return -m, -dm * dc

    def _acquisition_posterior(self, method, opt_type, normvar, jitter):
        """the PyMC model the reference rebuilds every BO iteration (gpmcmc.py:699-824): x priors + Potential(acquisition)."""
        def potential(x):
            f, g = self.acquisition_grad(x, method=method, opt_type=opt_type, normvar=normvar, jitter=jitter)
            return -f, -g
        return XPosterior(self.priors, potential)

    def __x_model_opt(self, post, opt_method, rng, **kwargs):
        """MAP (random start, gpmcmc.py:831-832 / :1169-1170) or MCMC (:844-853 / :1175-1188) over the x model."""
        maxeval = kwargs.pop('maxeval', 5000)
        kwargs.pop('progressbar', None)
        if opt_method == 'map':
            restarts = int(kwargs.pop('restarts', 1))
            kwargs.pop('seed', None)
            z0 = rng.standard_normal((restarts, post.space.P))
            zs, lps = find_map_multi(post, z0, maxeval=maxeval, **kwargs)
            lps = np.where(np.isfinite(lps), lps, -np.inf)
            mp = post.space.hypers_dict(zs[int(np.argmax(lps))])
            return mp, mp
        if opt_method not in ('mcmc_mean', 'mcmc_map'):
            raise Exception('method must be one of map, mcmc_map, or mcmc_mean')
        skw = {k: kwargs.pop(k) for k in ('draws', 'tune', 'chains', 'seed', 'target_accept', 'max_leapfrog',
                                          'path_length', 'init_jitter', 'sampler', 'max_treedepth') if k in kwargs}
        skw.setdefault('chains', 4)
        skw.setdefault('seed', int(rng.integers(2 ** 31)))
        data = hmc_sample(post, **skw)
        if opt_method == 'mcmc_mean':
            mp = self.mean_extract(data)
        else:
            mp = self.map_extract(data)
            try:
                z, _, _ = find_map(post, post.space.initial_z(mp), maxeval=maxeval)
                mp = post.space.hypers_dict(z)
            except (ValueError, FloatingPointError, np.linalg.LinAlgError):
                pass
        return data, mp

    def __acquisition_model_opt(self, opt_method, method, opt_type, normvar, jitter, rng, **kwargs):
        if _frozen_programs(self.yconrevs[0]) is None:
            raise Exception('Error: the output transform has no device program; use opt_method="predict" or "DE"')
        post = self._acquisition_posterior(method, opt_type, normvar, jitter)
        _, mp = self.__x_model_opt(post, opt_method, rng, **kwargs)
        return np.array([[float(mp[f'x{j}']) for j in range(self.nx)]])

    def __refine(self, optf, xsamp, lbs, ubs, acq=None):
        """polish of one candidate.  With ``acq`` (see :meth:`acquisition_grad`) this is the reference's own refine
        step -- find_MAP of the x model started at the candidate (gpmcmc.py:833-837) -- every evaluation one device
        call returning value and analytic gradient; otherwise bounded L-BFGS-B where each gradient is ONE batched
        predict of 2 nx + 1 points (central differences)."""
        from scipy.optimize import minimize
        if acq is not None and acq(xsamp) is not None:
            # device errors (GPError) propagate: only numerical trouble of the host optimiser falls through to the
            # finite-difference polish below
            try:
                def potential(x):
                    f, g = acq(x)
                    return -f, -g
                post = XPosterior(self.priors, potential)
                z0 = post.space.z_from_theta(np.clip(xsamp[0], lbs, ubs))
                f0 = post.logp_dlogp(z0[None, :], False)[0][0]
                z, f1, _ = find_map(post, z0, maxeval=200)
                if np.isfinite(f1) and f1 >= f0:
                    return post.space.theta_from_z(z[None, :])[0]
                return xsamp
            except (ValueError, FloatingPointError, np.linalg.LinAlgError):
                pass
        span = ubs - lbs

        def fg(x):
            h = 1e-6 * span
            pts = np.vstack([x[None, :], x[None, :] + np.diag(h), x[None, :] - np.diag(h)])
            pts = np.clip(pts, lbs, ubs)
            f = optf(pts)
            g = (f[1:1 + self.nx] - f[1 + self.nx:]) / (pts[1:1 + self.nx].diagonal() - pts[1 + self.nx:].diagonal())
            return f[0], g
        try:
            res = minimize(fg, xsamp[0], jac=True, method='L-BFGS-B', bounds=list(zip(lbs, ubs)),
                           options=dict(maxiter=50))
            if np.isfinite(res.fun) and res.fun <= optf(xsamp)[0]:
                return np.array([res.x])
        except (ValueError, FloatingPointError, np.linalg.LinAlgError):
            pass
        return xsamp

    # ---- Bayesian inverse problem --------------------------------------------------------------------
    def __gh_stats_inv(self, y, yv, deg=8):
        """variance of the converted observation by Gauss-Hermite quadrature (gpmcmc.py:573-585); as in the
        reference the value of the LAST observation is returned (its loop overwrites the result)."""
        xi, wi = np.polynomial.hermite.hermgauss(deg)
        yi = np.sqrt(2 * yv[-1, 0]) * xi + y[-1, 0]
        yir = self.yconrevs[0].con(yi)
        ym = np.sum(wi * yir) / np.sqrt(np.pi)
        return np.sum(wi * np.power(yir, 2)) / np.sqrt(np.pi) - ym ** 2

    def inverse_posterior(self, yobs, yvarobs=None, jitter=1e-6):
        """log posterior over the unknown input point given observations ``yobs`` [nobs,1] (optionally with
        variances ``yvarobs`` [nobs,1]): the model of gpmcmc.py:1049-1165 as an :class:`XPosterior`.  Replicated
        quirks: square roots of the variances go on the diagonal (:1138-1149); without ``yvarobs``
        the observation rows get no noise at all; the mean function is not subtracted from ``yobs`` (:1134);
        ``yder`` is taken at the raw outputs (:1152-1153)."""
        from .gp import GPEngine
        if self.hypers is None:
            raise Exception('Model must be fitted before running Bayesian optimisation')
        yobs = np.asarray(yobs, dtype=np.float64).reshape(-1, 1)
        nobs = len(yobs)
        yo_c = self.yconrevs[0].con(yobs[:, 0])
        noise_t = np.sqrt(self.hypers['gv'] + jitter) if self.noise else np.sqrt(jitter)
        noise_o = 0.0
        if yvarobs is not None:
            noise_o = np.sqrt(self.__gh_stats_inv(yobs, np.asarray(yvarobs, dtype=np.float64).reshape(-1, 1)))
        if nobs > 1 and not noise_o > 0.0:
            raise Exception('Error: several observations of one point need yvarobs (singular covariance otherwise)')
        # A = K_tt + noise_t I : an engine with a noise term and no jitter, "gv" = the value the reference adds
        eng = GPEngine(nx=self.nx, kerns=self.kerns, ops=self.ops, noise=True, jitter=0.0, device=self.device)
        eng.set_data(self.xc, self.yc[:, 0])
        space = ParamSpace(self.nx, self.nkern, True, has_alpha='RatQuad' in self.kerns)
        hyp = dict(self.hypers)
        hyp['gv'] = float(noise_t)
        th = space.theta_from_hypers(hyp)
        from .gp import check_info
        ll, _, info = eng.loglik_grad(th, want_grad=False)
        infos = (int(info[0]), int(eng.factorize(th)[0]))
        check_info(infos, 'inverse_posterior')
        if infos != (0, 0):
            raise Exception('Error: covariance matrix not positive definite')
        c = self.yconrevs[0]
        yfull = np.r_[self.y[:, 0], yobs[:, 0]]
        lyder = 0.0 if isinstance(c, _none_conrev) else float(np.sum(np.log(c.der(yfull))))
        cfull, cdiag = kdiag_values(self.kerns, self.ops, self.hypers['kv'], float(self.hypers.get('alpha', 1.0)))
        pot = InverseLikelihood(eng, self._con_with_der, yo_c, noise_o, cfull - cdiag, float(ll[0]) + lyder)
        return XPosterior(self.priors, pot)

    def inverse_opt(self, yobs, yvarobs=None, method='map', evaluate_opt=False, jitter=1e-6, seed=None, **kwargs):
        """Bayesian inverse solver (gpmcmc.py:1040-1217): the input point that explains ``yobs`` under the fitted
        surrogate, by MAP (random unconstrained start; ``restarts=R`` runs R starts as one device batch) or by MCMC
        over x.  Returns ``(data, xopt)`` or, with ``evaluate_opt``, ``(data, xopt, ysamp)`` after appending the
        evaluated point to the data set."""
        if self.verbose:
            print('Running Bayesian inverse solver...')
        post = self.inverse_posterior(yobs, yvarobs, jitter)
        data, mp = self.__x_model_opt(post, method, np.random.default_rng(seed), **kwargs)
        xopt = np.array([[float(mp[f'x{j}']) for j in range(self.nx)]])
        verb, self.verbose = self.verbose, False
        try:
            ypred = self.predict(xopt)
        finally:
            self.verbose = verb
        if self.verbose:
            print(f'Predicted {ypred} at x point {xopt}')
        if evaluate_opt:
            xsamp, ysamp = self._core__vector_solver(xopt)
            ym = self.__mean_values(xsamp)
            self.x = np.r_[self.x, xsamp]
            self.y = np.r_[self.y, ysamp]
            self.ym = np.r_[self.ym, ym]
            self.nsamp = len(self.x)
            self.__con(len(xsamp))
            if self.verbose:
                print(f'Actual evaluation is {ysamp} at x point {xsamp}')
            return data, xopt[0, :], ysamp[0]
        return data, xopt[0, :]

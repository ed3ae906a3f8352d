"""Latin-hypercube sampling and the (x, y) data holder (host side, thin).

Mirrors the reference's ``LHC`` (andvaranaut/lhc.py:17-131): ``sample`` (:24-37), ``__latin_sample`` (:40-47),
``del_samples`` (:50-93), ``set_data`` checks (:113-131).  Plot and netCDF helpers are not part of the GP hot path
and are left out (seaborn / netCDF4 are not in this stack).  Superset behaviour: the ``seed`` argument is honoured
(the reference accepts it and ignores it, lhc.py:40-43).
"""
import numpy as np
from scipy.stats import qmc

from .core import _core

__all__ = ['LHC']


class LHC(_core):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.x = np.empty((0, self.nx))
        self.y = np.empty((0, self.ny))

    def sample(self, nsamps, seed=None):
        if not isinstance(nsamps, int) or nsamps < 1:
            raise Exception('Error: nsamps argument must be an integer > 0')
        if self.verbose:
            print(f'Evaluating {nsamps} latin hypercube samples...')
        xsamps = self.__latin_sample(nsamps, seed)
        if self.constraints is not None:
            xsamps = self._core__check_constraints(xsamps)
        xsamps, ysamps = self._core__vector_solver(xsamps)
        self.x = np.r_[self.x, xsamps]
        self.y = np.r_[self.y, ysamps]
        self.nsamp = len(self.x)

    def __latin_sample(self, nsamps, seed=None):
        # random-cd optimisation is O(n^2) per sweep: keep it for the sizes the reference is used at
        opt = 'random-cd' if nsamps <= 2000 else None
        points = qmc.LatinHypercube(d=self.nx, optimization=opt, seed=seed).random(n=nsamps)
        xs = np.empty_like(points)
        for j in range(self.nx):
            xs[:, j] = self.priors[j].ppf(points[:, j])
        return xs

    def del_samples(self, ndels=None, method='coarse_lhc', idx=None):
        self.__del_samples(ndels, method, idx, returns=False)
        self.nsamp = len(self.x)

    def __del_samples(self, ndels, method, idx, returns):
        if method == 'coarse_lhc':
            if not isinstance(ndels, int) or ndels < 1:
                raise Exception('Error: must specify positive int for ndels')
            xs = self.__latin_sample(ndels)
            dmins = np.zeros(ndels, dtype=np.intc)
            for i in range(ndels):
                dmins[i] = int(np.argmin(np.linalg.norm(self.x - xs[i], axis=1)))
                self.x = np.delete(self.x, dmins[i], axis=0)
                self.y = np.delete(self.y, dmins[i], axis=0)
            return dmins if returns else None
        if method == 'random':
            if not isinstance(ndels, int) or ndels < 1:
                raise Exception('Error: must specify positive int for ndels')
            inds = np.random.choice(np.arange(len(self.x)), size=len(self.x) - ndels, replace=False)
            self.x, self.y = self.x[inds, :], self.y[inds, :]
            return inds if returns else None
        if method == 'specific':
            if not isinstance(idx, (int, list)):
                raise Exception('Error: must specify int or list of ints for idx')
            mask = np.ones(len(self.x), dtype=bool)
            mask[idx] = False
            self.x, self.y = self.x[mask], self.y[mask]
            return mask if returns else None
        raise Exception("Error: method must be one of 'coarse_lhc','random','specific'")

    def set_data(self, x, y):
        if not isinstance(x, np.ndarray) or x.ndim != 2 or x.dtype != 'float64' or x.shape[1] != self.nx:
            raise Exception('Error: Setting data requires a 2d numpy array of float64 inputs')
        if not isinstance(y, np.ndarray) or y.ndim != 2 or y.dtype != 'float64' or y.shape[1] != self.ny:
            raise Exception('Error: Setting data requires a 2d numpy array of float64 outputs')
        for i in range(self.nx):
            lo, hi = self.priors[i].interval(1.0)
            if not np.all(x[:, i] >= lo) or not np.all(x[:, i] <= hi):
                raise Exception('Error: provided x data must fit within provided input distribution ranges.')
        self.x, self.y = x, y
        self.nsamp = len(x)

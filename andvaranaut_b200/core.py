"""Target-function runner and persistence helpers (host side, thin).

Mirrors the constructor contract and the private solver of the reference's ``_core``
(andvaranaut/core.py:53-256): argument validation (:54-100), serial evaluation with failed / non-finite
samples dropped (:137-215), constraint filter (:218-246).  The reference's dask client (:105-134) is replaced
by a ``concurrent.futures`` process pool (dask is not part of this stack); target evaluation is outside the
GPU hot path either way.
"""
import multiprocessing as mp
from concurrent.futures import ProcessPoolExecutor
from time import time as stopwatch

import numpy as np

__all__ = ['_core', 'save_object', 'load_object']


def save_object(obj, fname):
    import cloudpickle
    with open(fname, 'wb') as f:
        cloudpickle.dump(obj, f)


def load_object(fname):
    import cloudpickle
    with open(fname, 'rb') as f:
        return cloudpickle.load(f)


def _is_scipy_frozen(p):
    return getattr(p, '__module__', None) == 'scipy.stats._distn_infrastructure'


def _run_one(args):
    fun, x = args
    return fun(x)


class _core:
    def __init__(self, nx, ny, priors, target, parallel=False, nproc=1, constraints=None, rundir=None,
                 verbose=True, pulse=1):
        if not isinstance(nx, int) or nx < 1:
            raise Exception('Error: must specify an integer number of input dimensions > 0')
        if not isinstance(ny, int) or ny < 1:
            raise Exception('Error: must specify an integer number of output dimensions > 0')
        if not isinstance(priors, list) or len(priors) != nx or not all(_is_scipy_frozen(p) for p in priors):
            raise Exception('Error: must provide list of scipy.stats univariate priors of length nx')
        if not callable(target):
            raise Exception('Error: must provide target function which produces output from specified inputs')
        if not isinstance(parallel, bool):
            raise Exception('Error: parallel must be type bool.')
        if not isinstance(nproc, int) or nproc < 1:
            raise Exception('Error: nproc argument must be an integer > 0')
        assert nproc <= mp.cpu_count(), 'Error: number of processors selected exceeds available.'
        keys = ['constraints', 'lower_bounds', 'upper_bounds']
        if constraints is not None and (not isinstance(constraints, dict) or not all(k in constraints for k in keys)):
            raise Exception(f'Error: provided constraints must be a dictionary with keys {keys} and list items.')
        self.nx, self.ny = nx, ny
        self.priors = priors
        self.target = target
        self.parallel = parallel
        self.nproc = nproc
        self.pulse = pulse
        self.constraints = constraints
        self.verbose = verbose
        self.rundir = rundir if rundir is not None else 'runs'
        self.nsamp = 0

    # evaluates ``fun`` (default: the target) at every row of xsamps; rows that raise or return
    # nan/inf are dropped from both arrays, as in the reference
    def __vector_solver(self, xsamps, fun=None):
        fun = self.target if fun is None else fun
        t0 = stopwatch()
        n = len(xsamps)
        outs, ok = [], np.ones(n, dtype=bool)
        if self.parallel and n > 1:
            with ProcessPoolExecutor(max_workers=self.nproc) as ex:
                futs = [ex.submit(_run_one, (fun, xsamps[i, :])) for i in range(n)]
                for i, f in enumerate(futs):
                    try:
                        outs.append(np.atleast_1d(f.result()))
                    except Exception as e:
                        print(f'Warning: Target function evaluation failed at sample {i} with x values: '
                              f'{xsamps[i, :]}; error message: {e}')
                        ok[i] = False
        else:
            for i in range(n):
                try:
                    outs.append(np.atleast_1d(fun(xsamps[i, :])))
                except Exception as e:
                    print(f'Warning: Target function evaluation failed at sample {i} with x values: '
                          f'{xsamps[i, :]}; error message: {e}')
                    ok[i] = False
        xs = xsamps[ok]
        if outs:
            try:
                ys = np.vstack(outs).astype(np.float64)
            except Exception:
                raise Exception('Error: number of target function outputs is not equal to ny')
            if ys.shape[1] != self.ny:
                raise Exception('Error: number of target function outputs is not equal to ny')
        else:
            ys = np.empty((0, self.ny))
        bad = ~np.all(np.isfinite(ys), axis=1)
        for i in np.where(bad)[0]:
            print(f'Warning: Target function evaluation returned inf/nan at sample with x values: {xs[i, :]}\n'
                  'Check range of input values valid.')
        xs, ys = xs[~bad], ys[~bad]
        if self.verbose:
            print(f'Time taken: {stopwatch() - t0:0.2f} s')
        return xs, ys

    def __check_constraints(self, xsamps):
        keep = np.ones(len(xsamps), dtype=bool)
        for i, x in enumerate(xsamps):
            for e, f in enumerate(self.constraints['constraints']):
                res = np.atleast_1d(f(x))
                lo = np.atleast_1d(self.constraints['lower_bounds'][e])
                hi = np.atleast_1d(self.constraints['upper_bounds'][e])
                if np.any(res < lo) or np.any(res > hi):
                    keep[i] = False
                    print(f'Sample {i + 1} with x values {x} removed due to invalidaing constraint {e + 1}.')
        if (~keep).any():
            print(f'{(~keep).sum()} samples removed due to violating constraints.')
        return xsamps[keep]

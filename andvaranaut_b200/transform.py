"""Conversion / reversion ("conrev") transforms of the GP surrogate -- host side.

Mirrors the public surface of the reference's ``andvaranaut/transform.py`` (same class names,
constructor arguments, ``con`` / ``rev`` / ``der`` semantics, ``wgp`` attributes ``warping_names,
params, np, pos, pid``) so user code switches over unchanged.  What is different by design:

* every learnable stage is described by one row of ``STAGES`` (opcode, parameter count, positivity
  flags) instead of a class per PyTensor expression; the reference's ``conmc/revmc/dermc`` symbolic
  twins (``transform.py:202-207,224-229,309-314,...,556-574``) do not exist here -- ``wgp.program()``
  emits the opcode list that the CUDA warp kernel (``csrc/warp.cuh``) evaluates together with its
  parameter Jacobians, and ``wgp.rev_program()`` the frozen inverse used by the predict epilogue;
* data-dependent stages (meanstd, stddev, stdshift, minshift, maxmin, pzero) freeze their affine
  coefficients at construction in NumPy mode exactly as the reference does (``transform.py:230-280,
  421-428``); on the device they are recomputed per hyperparameter sample.

Reference semantics followed: stage formulas ``transform.py:193-428``; composite ``wgp`` ordering,
parameter packing and running ``yzero`` ``transform.py:431-554``.
"""
import numpy as np

__all__ = ['normal', 'logit_logistic', 'probit', 'cdf', 'nonneg', 'log1p', 'log10', 'normalise',
           'quantile', 'robust', 'powerT', 'logarithm', 'affine', 'meanstd', 'minshift', 'stddev',
           'stdshift', 'maxmin', 'uniform', 'arcsinh', 'boxcox', 'boxcoxf', 'sinharcsinh', 'sal',
           'kumaraswamy', 'preserve_zero', 'wgp']

# opcodes shared with csrc/avn_types.h (keep in sync with enum avn_warp_op)
OP_AFFINE_CONST, OP_AFFINE, OP_LOG, OP_ARCSINH, OP_BOXCOX, OP_SINHARCSINH, OP_SAL, OP_KUMARASWAMY, \
    OP_STDSHIFT, OP_MEANSTD, OP_MINSHIFT, OP_STDDEV, OP_MAXMIN, OP_PZERO, OP_BOXCOX_CONST = range(15)

# name -> (opcode, positivity flags of its learnable parameters)
STAGES = {
    'affine': (OP_AFFINE, (False, True)),
    'logarithm': (OP_LOG, ()),
    'arcsinh': (OP_ARCSINH, (False, True, False, True)),
    'boxcox': (OP_BOXCOX, (False,)),
    'sinharcsinh': (OP_SINHARCSINH, (False, True)),
    'sal': (OP_SAL, (False, True, False, True)),
    'kumaraswamy': (OP_KUMARASWAMY, (True, True)),
    'stdshift': (OP_STDSHIFT, (False,)),
    'meanstd': (OP_MEANSTD, ()),
    'minshift': (OP_MINSHIFT, ()),
    'stddev': (OP_STDDEV, ()),
    'maxmin': (OP_MAXMIN, ()),
    'pzero': (OP_PZERO, ()),
    'uniform': (OP_AFFINE_CONST, ()),
    'boxcoxf': (OP_BOXCOX_CONST, ()),
}
_NEEDS_DATA = ('stdshift', 'meanstd', 'minshift', 'stddev', 'maxmin', 'pzero', 'boxcoxf')


# ---------------------------------------------------------------------------------------------
# fixed (non-learnable) conversions: applied once on the host, never on the device
# ---------------------------------------------------------------------------------------------
_LOGIT_BND = 0.9999999999999999
_LOGISTIC_BND = 36.7368005696771


def _clipped_logit(p):
    p = np.clip(p, 1.0 - _LOGIT_BND, _LOGIT_BND)
    return np.log(p) - np.log1p(-p)


def _clipped_logistic(x):
    x = np.clip(x, -_LOGISTIC_BND, _LOGISTIC_BND)
    s = np.sign(x)
    ex = np.exp(s * x)
    return 0.5 - s * 0.5 + s * ex / (ex + 1.0)


def _two_sided_cdf(x, dist, pivot):
    # evaluate through the survival function on the lower side, as the reference does
    return np.where(x < pivot, 1 - dist.sf(x), dist.cdf(x))


def _two_sided_ppf(p, dist):
    return np.where(p < 0.5, dist.isf(1 - p), dist.ppf(p))


class normal:
    """standardise by the prior's mean/std (reference ``normal``, transform.py:139-142)."""

    def __init__(self, dist):
        self._m, self._s = dist.mean(), dist.std()

    def con(self, x):
        return (x - self._m) / self._s

    def rev(self, x):
        return x * self._s + self._m


class cdf:
    """prior CDF to the unit interval (transform.py:151-154)."""

    def __init__(self, dist):
        self._d = dist

    def con(self, x):
        return _two_sided_cdf(x, self._d, self._d.mean())

    def rev(self, x):
        return _two_sided_ppf(x, self._d)


class logit_logistic(cdf):
    """prior CDF followed by a clipped logit (transform.py:143-146)."""

    def con(self, x):
        return _clipped_logit(super().con(x))

    def rev(self, x):
        return super().rev(_clipped_logistic(x))


class probit:
    """prior CDF followed by the standard-normal quantile (transform.py:147-150)."""

    def __init__(self, dist):
        import scipy.stats as st
        self._d, self._n = dist, st.norm()

    def con(self, x):
        return _two_sided_ppf(_two_sided_cdf(x, self._d, 0), self._n)

    def rev(self, x):
        return _two_sided_ppf(_two_sided_cdf(x, self._n, 0), self._d)


class nonneg:
    def con(self, y):
        return _clipped_logit(y / (1 + y))

    def rev(self, y):
        p = _clipped_logistic(y)
        return p / (1 - p)


class log1p:
    con = staticmethod(np.log1p)
    rev = staticmethod(np.expm1)


class log10:
    con = staticmethod(np.log10)

    @staticmethod
    def rev(y):
        return np.power(10, y)


class normalise:
    def __init__(self, fac):
        self.fac = fac

    def con(self, y):
        return y / self.fac

    def rev(self, y):
        return y * self.fac


class _sk:
    """thin adaptor over a fitted scikit-learn transformer (transform.py:171-192)."""

    def con(self, y):
        return self._t.transform(np.reshape(y, (-1, 1)))[:, 0]

    def rev(self, y):
        return self._t.inverse_transform(np.reshape(y, (-1, 1)))[:, 0]


class quantile(_sk):
    def __init__(self, x, mode='normal'):
        from sklearn.preprocessing import QuantileTransformer
        self.mode = mode
        self._t = self.qt = QuantileTransformer(output_distribution=mode).fit(np.reshape(x, (-1, 1)))


class robust(_sk):
    def __init__(self, x):
        from sklearn.preprocessing import RobustScaler
        self._t = self.rs = RobustScaler().fit(np.reshape(x, (-1, 1)))


class powerT(_sk):
    def __init__(self, x, method='yeo-johnson'):
        from sklearn.preprocessing import PowerTransformer
        self.method = method
        self._t = self.pt = PowerTransformer(method=method).fit(np.reshape(x, (-1, 1)))
        self.pt.lambdas_[0] = np.minimum(np.maximum(-0.01, self.pt.lambdas_[0]), 1.0)


# ---------------------------------------------------------------------------------------------
# differentiable stages (usable inside wgp and on the device)
# ---------------------------------------------------------------------------------------------
class _stage:
    op = None

    def coeffs(self):
        """constants handed to the device when the stage is frozen (<= 4 doubles)."""
        return ()


class logarithm(_stage):
    op = OP_LOG

    def con(self, y):
        return np.log(y)

    def rev(self, y):
        return np.exp(y)

    def der(self, y):
        return 1 / y


class affine(_stage):
    """a + b*y; base of every data-dependent normalisation."""
    op = OP_AFFINE

    def __init__(self, a, b):
        self.a, self.b = a, b

    def con(self, y):
        return self.a + self.b * y

    def rev(self, y):
        return (y - self.a) / self.b

    def der(self, y):
        return self.b * np.ones_like(y)

    def coeffs(self):
        return (float(self.a), float(self.b))


class meanstd(affine):
    def __init__(self, y):
        m, s = np.mean(y), np.std(y)
        affine.__init__(self, -m / s, 1 / s)


class minshift(affine):
    def __init__(self, y, safety=1000):
        affine.__init__(self, -np.min(y) * safety, 1.0)


class stddev(affine):
    def __init__(self, y):
        affine.__init__(self, 0, 1 / np.std(y))


class stdshift(affine):
    def __init__(self, a, y):
        affine.__init__(self, a, 1 / np.std(y))


class maxmin(affine):
    def __init__(self, x, centred=False, safety=0.01):
        lo, hi = np.min(x), np.max(x)
        span = (hi - lo) / (1 - 2 * safety)
        if centred:
            affine.__init__(self, -(hi + lo) / span, 2 / span)
        else:
            affine.__init__(self, -lo / span + safety, 1 / span)


class uniform(affine):
    def __init__(self, dist, safety=1e-10):
        lo, hi = dist.interval(1.0)
        span = (hi - lo) / (1 - 2 * safety)
        affine.__init__(self, -lo / span + safety, 1 / span)


class preserve_zero(affine):
    def __init__(self, y, yzero):
        s = np.std(y)
        affine.__init__(self, -yzero / s, 1 / s)


class arcsinh(_stage):
    op = OP_ARCSINH

    def __init__(self, a, b, c, d):
        self.a, self.b, self.c, self.d = a, b, c, d

    def con(self, y):
        return self.a + self.b * np.arcsinh((y - self.c) / self.d)

    def rev(self, y):
        return self.c + self.d * np.sinh((y - self.a) / self.b)

    def der(self, y):
        return self.b / np.sqrt(self.d ** 2 + (y - self.c) ** 2)

    def coeffs(self):
        return (float(self.a), float(self.b), float(self.c), float(self.d))


class boxcox(_stage):
    """sign-preserving Box-Cox with exponent lamb+1 (identity at lamb=0)."""
    op = OP_BOXCOX

    def __init__(self, lamb):
        self.lamb = lamb

    def con(self, y):
        q = self.lamb + 1
        return (np.sign(y) * np.abs(y) ** q - 1) / q

    def rev(self, y):
        q = self.lamb + 1
        t = y * q + 1
        return np.sign(t) * np.abs(t) ** (1 / q)

    def der(self, y):
        return np.abs(y) ** self.lamb

    def coeffs(self):
        return (float(self.lamb),)


class boxcoxf(boxcox):
    def __init__(self, y):
        from sklearn.preprocessing import PowerTransformer
        powt = PowerTransformer(method='box-cox', standardize=False).fit(np.reshape(y, (-1, 1)))
        self.lamb = powt.lambdas_[0]


class sinharcsinh(_stage):
    op = OP_SINHARCSINH

    def __init__(self, a, b):
        self.a, self.b = a, b

    def _u(self, y):
        return self.b * np.arcsinh(y) - self.a

    def con(self, y):
        return np.sinh(self._u(y))

    def rev(self, y):
        return np.sinh((np.arcsinh(y) + self.a) / self.b)

    def der(self, y):
        return self.b * np.cosh(self._u(y)) / np.sqrt(1 + y ** 2)

    def coeffs(self):
        return (float(self.a), float(self.b))


class sal(sinharcsinh):
    """sinh-arcsinh followed by an affine map c + d*(.)."""
    op = OP_SAL

    def __init__(self, a, b, c, d):
        self.a, self.b, self.c, self.d = a, b, c, d

    def con(self, y):
        return self.c + self.d * np.sinh(self._u(y))

    def rev(self, y):
        return np.sinh((np.arcsinh((y - self.c) / self.d) + self.a) / self.b)

    def der(self, y):
        return self.d * sinharcsinh.der(self, y)

    def coeffs(self):
        return (float(self.a), float(self.b), float(self.c), float(self.d))


class kumaraswamy(_stage):
    """Kumaraswamy CDF on [0,1] (input warping)."""
    op = OP_KUMARASWAMY

    def __init__(self, a, b):
        self.a, self.b = a, b

    def con(self, x):
        return 1 - (1 - x ** self.a) ** self.b

    def rev(self, x):
        return (1 - (1 - x) ** (1 / self.b)) ** (1 / self.a)

    def der(self, x):
        return self.a * self.b * x ** (self.a - 1) * (1 - x ** self.a) ** (self.b - 1)

    def coeffs(self):
        return (float(self.a), float(self.b))


# ---------------------------------------------------------------------------------------------
# composite warp
# ---------------------------------------------------------------------------------------------
class wgp:
    """Chain of stages with one flat parameter vector (reference ``wgp``, transform.py:431-554).

    ``warpings``  list of stage names; ``params`` flat vector consumed left to right;
    ``y`` the data the data-dependent stages take their statistics from (pushed through the chain
    as it is built); ``xdist`` the scipy prior needed by ``'uniform'``.
    """

    def __init__(self, warpings, params, y=None, xdist=None, mode='numpy'):
        if mode != 'numpy':
            raise ValueError("only mode='numpy' exists on the host; the symbolic mode of the reference "
                             "is replaced by wgp.program() evaluated on the device")
        self.warping_names = list(warpings)
        self.params = np.asarray(params, dtype=np.float64).reshape(-1)
        self.xdist = xdist
        self.warpings = []
        self.pid = np.zeros(len(self.warping_names), dtype=np.int32)
        self.pos = np.zeros(len(self.params), dtype=np.bool_)
        run = None if y is None else np.array(y, dtype=np.float64, copy=True)
        zero = 0.0
        k = 0
        for s, name in enumerate(self.warping_names):
            if name not in STAGES:
                raise Exception(f'Only {sorted(STAGES)} classes allowed')
            if name in _NEEDS_DATA and run is None:
                raise Exception(f'Must supply y array to use {name}')
            flags = STAGES[name][1]
            p = self.params[k:k + len(flags)]
            if len(p) != len(flags):
                raise Exception(f'wgp: not enough parameters for stage {name}')
            self.pos[k:k + len(flags)] = flags
            if name == 'uniform':
                if xdist is None:
                    raise Exception('Must supply x distribution to use uniform')
                st = uniform(xdist)
            elif name == 'pzero':
                st = preserve_zero(run, zero)
            elif name == 'stdshift':
                st = stdshift(p[0], run)
            elif name in ('meanstd', 'minshift', 'stddev', 'maxmin', 'boxcoxf'):
                st = globals()[name](run)
            else:
                st = globals()[name](*p)
            st.name = name
            self.warpings.append(st)
            k += len(flags)
            self.pid[s] = k
            if run is not None:
                with np.errstate(divide='ignore', invalid='ignore'):
                    run = st.con(run)
                    zero = st.con(zero)
        self.np = k

    def con(self, y):
        for st in self.warpings:
            y = st.con(y)
        return y

    def rev(self, y):
        for st in reversed(self.warpings):
            y = st.rev(y)
        return y

    def der(self, y):
        g = np.ones_like(y, dtype=np.float64)
        for st in self.warpings:
            g = g * st.der(y)
            y = st.con(y)
        return g

    # -- device programs -------------------------------------------------------------------
    def program(self):
        """Learnable program: list of (opcode, first-parameter index or -1, 4 constants).  The
        data-dependent stages keep their opcode so the device recomputes the statistics for every
        hyperparameter sample (the reference's PyTensor mode)."""
        prog = []
        k = 0
        for st in self.warpings:
            op, flags = STAGES[st.name]
            c = [0.0] * 4
            if op in (OP_AFFINE_CONST, OP_BOXCOX_CONST):
                cc = st.coeffs()
                c[:len(cc)] = cc
            prog.append((op, k if len(flags) else -1, tuple(c)))
            k += len(flags)
        return prog

    def rev_program(self):
        """Frozen program (all coefficients constant), evaluated right-to-left by the predict
        epilogue for the Gauss-Hermite reversion (gpmcmc.py:551)."""
        return [frozen_stage(st) for st in self.warpings]


def frozen_stage(st):
    """(opcode, -1, constants) of one differentiable stage with its coefficients frozen; None if the object is
    not one of the device-evaluable stages."""
    if not isinstance(st, _stage):
        return None
    if isinstance(st, affine):
        op = OP_AFFINE_CONST
    elif isinstance(st, boxcox):
        op = OP_BOXCOX_CONST
    else:
        op = st.op
    c = [0.0] * 4
    cc = st.coeffs()
    c[:len(cc)] = cc
    return (op, -1, tuple(c))

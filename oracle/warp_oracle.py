"""Oracle restatement of andvaranaut's composite warp ``wgp`` (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/andvaranaut/transform.py``:
  * stage formulas ``con/rev/der``: affine :208-229, meanstd :230-239, minshift :240-247,
    stddev :248-255, stdshift :256-264, maxmin :265-280, uniform :281-288, arcsinh :289-314,
    boxcox :316-337, sinharcsinh :344-365, sal :366-391, kumaraswamy :394-417,
    preserve_zero :421-428, logarithm :193-207;
  * composition order, parameter packing (``pid``/``pos``/``np``), running data ``yc`` and
    running zero image ``yzero``: ``wgp.__init__`` :431-534; ``con`` :536-540, ``rev`` :542-546,
    ``der`` :548-554.

The reference evaluates the chain twice: in NumPy (frozen statistics) and symbolically in
PyTensor, where the data-dependent stages (meanstd, stddev, stdshift, minshift, maxmin, pzero)
are functions of the preceding learnable parameters and are differentiated by autodiff
(``transform.py:448-452,527-533``; used at ``gpmcmc.py:224-231,275-277``).  Here the symbolic
evaluation is restated with forward-mode dual numbers carried through exactly the same
expressions, so values and parameter-derivatives come from one code path and the CUDA kernels
(hand-derived per-stage Jacobians) are checked against an independent derivation.
"""
import numpy as np

ALLOWED = ['affine', 'logarithm', 'arcsinh', 'boxcox', 'sinharcsinh', 'sal',
           'meanstd', 'boxcoxf', 'uniform', 'maxmin', 'kumaraswamy', 'pzero',
           'stddev', 'stdshift', 'minshift']

# number of learnable parameters and positivity flags per stage (transform.py:458-496)
STAGE_PARAMS = {
    'affine': (False, True),
    'logarithm': (),
    'arcsinh': (False, True, False, True),
    'boxcox': (False,),
    'sinharcsinh': (False, True),
    'sal': (False, True, False, True),
    'kumaraswamy': (True, True),
    'stdshift': (False,),
    'meanstd': (), 'minshift': (), 'stddev': (), 'maxmin': (), 'pzero': (), 'uniform': (),
}


class Dual:
    """value ``v`` (any shape S) with derivatives ``d`` of shape S + (P,)."""
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v = np.asarray(v, dtype=np.float64)
        self.d = np.asarray(d, dtype=np.float64)

    @staticmethod
    def lift(x, P):
        if isinstance(x, Dual):
            return x
        x = np.asarray(x, dtype=np.float64)
        return Dual(x, np.zeros(x.shape + (P,)))

    @property
    def P(self):
        return self.d.shape[-1]

    def _b(self, o):
        return Dual.lift(o, self.P)

    def __add__(self, o):
        o = self._b(o)
        return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._b(o)
        return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return self._b(o) - self

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __mul__(self, o):
        o = self._b(o)
        return Dual(self.v * o.v, self.d * o.v[..., None] + o.d * self.v[..., None])
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._b(o)
        q = self.v / o.v
        return Dual(q, (self.d - o.d * q[..., None]) / o.v[..., None])

    def __rtruediv__(self, o):
        return self._b(o) / self


def _un(f, df):
    def g(x):
        if isinstance(x, Dual):
            return Dual(f(x.v), x.d * df(x.v)[..., None])
        return f(x)
    return g


dlog = _un(np.log, lambda v: 1.0 / v)
dexp = _un(np.exp, np.exp)
dsinh = _un(np.sinh, np.cosh)
dcosh = _un(np.cosh, np.sinh)
darcsinh = _un(np.arcsinh, lambda v: 1.0 / np.sqrt(1.0 + v * v))
dsqrt = _un(np.sqrt, lambda v: 0.5 / np.sqrt(v))
dabs = _un(np.abs, np.sign)


def dsign(x):
    return np.sign(x.v) if isinstance(x, Dual) else np.sign(x)


def dpow(x, p):
    """x**p for x > 0 (or x == 0 with constant exponent handled by numpy), p scalar Dual/float."""
    if not isinstance(x, Dual) and not isinstance(p, Dual):
        return np.power(x, p)
    P = x.P if isinstance(x, Dual) else p.P
    x = Dual.lift(x, P)
    p = Dual.lift(p, P)
    val = np.power(x.v, p.v)
    with np.errstate(divide='ignore', invalid='ignore'):
        dx = p.v * np.power(x.v, p.v - 1.0)
        lx = np.where(x.v > 0, np.log(np.where(x.v > 0, x.v, 1.0)), 0.0)
    d = x.d * dx[..., None] + p.d * (val * lx)[..., None]
    return Dual(val, d)


def dmean(x):
    if isinstance(x, Dual):
        return Dual(np.mean(x.v), np.mean(x.d, axis=0))
    return np.mean(x)


def dstd(x):
    """population standard deviation (ddof=0), as np.std / pt.std."""
    if isinstance(x, Dual):
        m = np.mean(x.v)
        s = np.std(x.v)
        dm = np.mean(x.d, axis=0)
        ds = np.sum((x.v - m)[:, None] * (x.d - dm[None, :]), axis=0) / (x.v.size * s)
        return Dual(s, ds)
    return np.std(x)


def dmin(x):
    if isinstance(x, Dual):
        i = int(np.argmin(x.v))
        return Dual(x.v[i], x.d[i])
    return np.min(x)


def dmax(x):
    if isinstance(x, Dual):
        i = int(np.argmax(x.v))
        return Dual(x.v[i], x.d[i])
    return np.max(x)


def _val(x):
    return x.v if isinstance(x, Dual) else x


# ---------------------------------------------------------------------------------------------
# stages: each is (con, rev, der) closures over its (possibly Dual) coefficients
# ---------------------------------------------------------------------------------------------
class _Stage:
    def __init__(self, name, con, rev, der, coeffs=None):
        self.name, self.con, self.rev, self.der = name, con, rev, der
        self.coeffs = coeffs or {}


def _affine_stage(name, a, b):
    def con(y):
        return a + b * y

    def rev(z):
        return (z - a) / b

    def der(y):
        one = np.ones_like(_val(y))
        return b * one
    return _Stage(name, con, rev, der, {'a': a, 'b': b})


def _make_stage(name, p, yc, yzero, xdist_interval, consts):
    """Build one stage from its parameters ``p`` (list of Dual/float) and running data ``yc``."""
    if name == 'affine':
        return _affine_stage(name, p[0], p[1])
    if name == 'logarithm':
        return _Stage(name, dlog, dexp, lambda y: 1.0 / y)
    if name == 'arcsinh':
        a, b, c, d = p
        return _Stage(name,
                      lambda y: a + b * darcsinh((y - c) / d),
                      lambda z: c + d * dsinh((z - a) / b),
                      lambda y: b / dsqrt(d * d + (y - c) * (y - c)))
    if name in ('boxcox', 'boxcoxf'):
        lamb = p[0] if name == 'boxcox' else consts['lamb']
        lambp = lamb + 1

        def con(y):
            return (dsign(y) * dpow(dabs(y), lambp) - 1) / lambp

        def rev(z):
            term = z * lambp + 1
            return dsign(term) * dpow(dabs(term), 1 / lambp)

        def der(y):
            return dpow(dabs(y), lamb)
        return _Stage(name, con, rev, der, {'lamb': lamb})
    if name == 'sinharcsinh':
        a, b = p
        return _Stage(name,
                      lambda y: dsinh(b * darcsinh(y) - a),
                      lambda z: dsinh((darcsinh(z) + a) / b),
                      lambda y: b * dcosh(b * darcsinh(y) - a) / dsqrt(1 + y * y))
    if name == 'sal':
        a, b, c, d = p
        return _Stage(name,
                      lambda y: c + d * dsinh(b * darcsinh(y) - a),
                      lambda z: dsinh((darcsinh((z - c) / d) + a) / b),
                      lambda y: b * d * dcosh(b * darcsinh(y) - a) / dsqrt(1 + y * y))
    if name == 'kumaraswamy':
        a, b = p
        return _Stage(name,
                      lambda x: 1 - dpow(1 - dpow(x, a), b),
                      lambda z: dpow(1 - dpow(1 - z, 1 / b), 1 / a),
                      lambda x: a * b * dpow(x, a - 1) * dpow(1 - dpow(x, a), b - 1))
    if name == 'stdshift':
        std = dstd(yc)
        return _affine_stage(name, p[0], 1 / std)
    if name == 'meanstd':
        mean, std = dmean(yc), dstd(yc)
        return _affine_stage(name, -mean / std, 1 / std)
    if name == 'minshift':
        mini = dmin(yc)
        return _affine_stage(name, -mini * 1000, 1.0)
    if name == 'stddev':
        std = dstd(yc)
        return _affine_stage(name, 0, 1 / std)
    if name == 'maxmin':
        safety = 0.01
        xmin, xmax = dmin(yc), dmax(yc)
        xminus = (xmax - xmin) / (1 - 2 * safety)
        return _affine_stage(name, -xmin / xminus + safety, 1 / xminus)
    if name == 'pzero':
        ystd = dstd(yc)
        return _affine_stage(name, -yzero / ystd, 1 / ystd)
    if name == 'uniform':
        safety = 1e-10
        lo, hi = xdist_interval
        xminus = (hi - lo) / (1 - 2 * safety)
        return _affine_stage(name, -lo / xminus + safety, 1 / xminus)
    raise ValueError(f'Only {ALLOWED} classes allowed')


class WarpOracle:
    """Restates ``wgp(warpings, params, y, xdist)``.

    ``params`` may be floats (NumPy mode of the reference) or, with ``with_duals=True``, they are
    seeded as independent dual variables so that every output carries d/dparams (the PyTensor
    mode of the reference followed by autodiff).
    """

    def __init__(self, warpings, params, y=None, xdist_interval=None, with_duals=False, consts=None):
        self.warping_names = list(warpings)
        params = np.asarray(params, dtype=np.float64).reshape(-1)
        P = len(params)
        self.params = params
        self.with_duals = with_duals
        if with_duals:
            eye = np.eye(P)
            pv = [Dual(params[i], eye[i]) for i in range(P)]
        else:
            pv = list(params)
        self.pos = np.zeros(P, dtype=bool)
        self.pid = np.zeros(len(warpings), dtype=np.int32)
        self.stages = []
        pc = 0
        yzero = 0.0
        yc = None
        if y is not None:
            yv = np.asarray(y, dtype=np.float64)
            yc = Dual.lift(yv, P) if with_duals else yv.copy()
            if with_duals:
                yzero = Dual.lift(0.0, P)
        consts = consts or {}
        for s, name in enumerate(warpings):
            if name not in ALLOWED:
                raise ValueError(f'Only {ALLOWED} classes allowed')
            flags = STAGE_PARAMS.get(name, ())
            npar = len(flags)
            needs_y = name in ('stdshift', 'meanstd', 'minshift', 'stddev', 'maxmin', 'pzero', 'boxcoxf')
            if needs_y and y is None:
                raise ValueError(f'Must supply y array to use {name}')
            if name == 'uniform' and xdist_interval is None:
                raise ValueError('Must supply x distribution to use uniform')
            st = _make_stage(name, pv[pc:pc + npar], yc, yzero, xdist_interval, consts.get(s, {}))
            self.pos[pc:pc + npar] = flags
            pc += npar
            self.pid[s] = pc
            self.stages.append(st)
            if y is not None:
                with np.errstate(divide='ignore', invalid='ignore'):
                    yc = st.con(yc)
                    yzero = st.con(yzero)
        self.np = pc
        self._ycon = yc   # con(y) of the construction data (Dual when with_duals)

    def con(self, y):
        res = y
        for st in self.stages:
            res = st.con(res)
        return res

    def rev(self, z):
        res = z
        for st in reversed(self.stages):
            res = st.rev(res)
        return res

    def der(self, y):
        res = np.ones_like(_val(y)) if not self.with_duals else Dual.lift(np.ones_like(_val(y)), len(self.params))
        x = y
        for st in self.stages:
            res = res * st.der(x)
            x = st.con(x)
        return res

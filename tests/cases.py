"""Shared seeded GP cases for the parity tests (inputs identical for the oracle and the CUDA path)."""
import os
import sys

import numpy as np
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))

from oracle.gp_oracle import ModelSpec  # noqa: E402
from andvaranaut_b200 import transform as T  # noqa: E402


def engine_args(spec: ModelSpec):
    """translate an oracle ModelSpec into GPEngine constructor arguments."""
    xw = None
    if spec.xwarps is not None:
        xw = []
        for w in spec.xwarps:
            if w is None:
                xw.append(None)
            else:
                names, interval = w
                xd = st.uniform(interval[0], interval[1] - interval[0]) if interval is not None else None
                npar = sum(len(T.STAGES[s][1]) for s in names)
                # program() only needs stage structure; dummy params and data
                prog = T.wgp(names, np.ones(npar), y=np.linspace(0.1, 0.9, 8), xdist=xd).program()
                xw.append(prog)
    yw = None
    if spec.ywarp is not None:
        npar = sum(len(T.STAGES[s][1]) for s in spec.ywarp)
        yw = T.wgp(spec.ywarp, np.ones(npar), y=np.linspace(0.5, 1.5, 8)).program()
    return dict(nx=spec.nx, kerns=spec.kerns, ops=spec.ops, noise=spec.noise, jitter=spec.jitter, xwarps=xw, ywarp=yw)


def synth(spec, N, seed, M=0, ywarp_positive=True):
    """seeded synthetic data + a plausible theta for any spec."""
    rng = np.random.default_rng(seed)
    d = spec.nx
    X = st.qmc.LatinHypercube(d=d, seed=seed).random(N)
    a = np.linspace(0.5, 2.0, d)
    y = np.exp(np.sum(np.sin(2 * np.pi * a * X), axis=1) / d + 0.5 * X[:, 0] * X[:, -1]) + 0.01 * rng.normal(size=N)
    if spec.ywarp is None:
        y = (y - y.mean()) / y.std()
    o = spec.offsets()
    th = np.zeros(o['P'])
    if spec.noise:
        th[o['gv']] = 1e-3 * np.exp(0.3 * rng.normal())
    th[o['l']:o['l'] + d * spec.nkern] = np.exp(0.3 * rng.normal(size=d * spec.nkern))
    th[o['kv']:o['kv'] + spec.nkern] = 1.5 * np.exp(0.3 * rng.normal(size=spec.nkern))
    th[o['iw']:o['iw'] + spec.n_iw()] = np.exp(0.2 * rng.normal(size=spec.n_iw()))
    if spec.ywarp is not None:
        from oracle.warp_oracle import STAGE_PARAMS
        vals = []
        for s in spec.ywarp:
            for pos in STAGE_PARAMS.get(s, ()):
                vals.append(np.exp(0.1 * rng.normal()) if pos else 0.1 * rng.normal())
        th[o['cw']:o['cw'] + len(vals)] = vals
    if spec.has_alpha:
        th[o['alpha']] = 1.7
    Xs = rng.uniform(0, 1, (M, d)) if M else np.zeros((0, d))
    return X, y, th, Xs

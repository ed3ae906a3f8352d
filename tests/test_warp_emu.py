"""CPU test of the device warp code (csrc/warp.cuh) compiled for the host with a one-thread block:
values, parameter Jacobians, log-derivative sums and the frozen inverse against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'golden'))

from andvaranaut_b200 import _lib, transform as T  # noqa: E402
from oracle.warp_oracle import WarpOracle, Dual  # noqa: E402
import make_golden as mg  # noqa: E402


@pytest.fixture(scope='module')
def emu():
    lib = os.path.join(HERE, 'host_emu', 'libwarp_emu.so')
    src = os.path.join(HERE, 'host_emu', 'warp_emu.cu')
    dep = os.path.join(ROOT, 'andvaranaut_b200', 'csrc', 'warp.cuh')
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(['g++', '-x', 'c++', '-O1', '-shared', '-fPIC', '-o', lib, src])
    e = C.CDLL(lib)
    e.warp_emu_run.argtypes = [C.POINTER(_lib.WarpProg), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_void_p]
    e.warp_emu_rev.argtypes = [C.POINTER(_lib.WarpProg), C.c_double]
    e.warp_emu_rev.restype = C.c_double
    return e


@pytest.mark.parametrize('case', mg.WGP_CASES, ids=[c[0] for c in mg.WGP_CASES])
def test_device_warp_math(emu, case):
    name, stages, params, kind, interval = case
    rng = np.random.default_rng(3)
    d = mg.data_of(kind, rng, 50)
    xd = st.uniform(interval[0], interval[1] - interval[0]) if interval else None
    params = np.array(params, dtype=np.float64)
    w = T.wgp(stages, params, y=d, xdist=xd)
    prog = _lib.make_prog(w.program())
    assert prog.nparams == len(params)
    val = d.copy()
    dual = np.zeros((len(d), 8))
    ls = np.zeros(1)
    dls = np.zeros(8)
    emu.warp_emu_run(C.byref(prog), params.ctypes.data, len(d), val.ctypes.data, dual.ctypes.data, 1,
                     ls.ctypes.data, dls.ctypes.data)
    wo = WarpOracle(stages, params, y=d, xdist_interval=interval, with_duals=True)
    z = wo._ycon
    der = wo.der(Dual.lift(d, len(params)))
    n = len(params)
    assert np.max(np.abs(val - z.v) / np.maximum(1e-12, np.abs(z.v))) < 1e-13
    if n:
        assert np.max(np.abs(dual[:, :n] - z.d)) <= 1e-13 * max(1.0, np.max(np.abs(z.d)))
        dl = np.sum(der.d / der.v[:, None], axis=0)
        assert np.max(np.abs(dls[:n] - dl)) <= 1e-12 * max(1.0, np.max(np.abs(dl)))
    assert abs(ls[0] - np.sum(np.log(der.v))) <= 1e-13 * max(1.0, abs(np.sum(np.log(der.v))))
    # frozen inverse program used by the predict epilogue
    rprog = _lib.make_prog(w.rev_program(), nparams=0)
    zz = w.con(d)
    back = np.array([emu.warp_emu_rev(C.byref(rprog), float(v)) for v in zz])
    assert np.max(np.abs(back - w.rev(zz)) / np.maximum(1e-12, np.abs(w.rev(zz)))) < 1e-12

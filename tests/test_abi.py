"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/avn_gp.h declares,
and its host-side argument checking works without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from andvaranaut_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'avn_gp.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(avn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/avn_gp.h but not exported'
    assert sorted(_lib.SYMBOLS) == names, 'ctypes prototypes out of sync with the header'


def test_struct_sizes_match_header():
    assert C.sizeof(_lib.WarpStage) == 40
    assert C.sizeof(_lib.WarpProg) == 8 + 6 * 40
    assert C.sizeof(_lib.ModelDesc) == 48 + 8 + 17 * 248
    assert C.sizeof(_lib.WsLayout) == 18 * 8


def test_create_and_parameter_layout_without_gpu():
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.d, d.nkern, d.noise, d.jitter = 3, 2, 1, 1e-6
    d.kern[0], d.kern[1], d.op[0] = 0, 4, 1
    d.xwarp[1] = _lib.make_prog([(0, -1, (0.0, 1.0, 0, 0)), (7, 0, (0, 0, 0, 0))])
    d.ywarp = _lib.make_prog([(2, -1, (0,) * 4), (6, 0, (0,) * 4), (9, -1, (0,) * 4)])
    h = C.c_void_p()
    assert lib.avn_gp_create(C.byref(d), C.byref(h)) == 0
    # gv + l(3*2) + kv(2) + iw(2) + cw(4) + alpha
    assert lib.avn_gp_num_params(h) == 1 + 6 + 2 + 2 + 4 + 1
    assert lib.avn_gp_workspace_bytes(h, 4) == 0          # no data yet
    lib.avn_gp_destroy(h)


@pytest.mark.parametrize('mutate,msg', [
    (lambda d: setattr(d, 'd', 0), 'd out of range'),
    (lambda d: setattr(d, 'd', 17), 'd out of range'),
    (lambda d: setattr(d, 'nkern', 5), 'nkern out of range'),
    (lambda d: d.kern.__setitem__(0, 9), 'unknown kernel'),
])
def test_create_rejects_bad_descriptions(mutate, msg):
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.d, d.nkern, d.noise, d.jitter = 2, 1, 1, 1e-6
    mutate(d)
    h = C.c_void_p()
    assert lib.avn_gp_create(C.byref(d), C.byref(h)) != 0
    assert msg in _lib.last_error()


def test_two_ratquad_rejected_like_reference():
    # gpmcmc.py:287 "Only works of only one ratquad kernel specified"
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.d, d.nkern, d.noise, d.jitter = 2, 2, 1, 1e-6
    d.kern[0] = d.kern[1] = 4
    h = C.c_void_p()
    assert lib.avn_gp_create(C.byref(d), C.byref(h)) != 0


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from andvaranaut_b200.gp import GPEngine, GPError
    with pytest.raises(GPError):
        GPEngine(nx=2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'andvaranaut_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                assert not re.search(r'^\s*(from|import)\s+oracle', open(os.path.join(dirpath, f)).read(), flags=re.M), f

"""ctypes binding of ``libavn_gp.so`` (the C ABI declared in ``include/avn_gp.h``).

There is no CPU fallback: if the shared library is missing the import of the GP path fails loudly
with the build command to run.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVN_GP_LIB selects an instrumented build of the same library (e.g. -DAVN_FACTOR_PROF); default: the in-tree one
LIB_PATH = os.environ.get('AVN_GP_LIB') or os.path.join(_HERE, 'libavn_gp.so')

AVN_MAX_D = 16
AVN_MAX_KERN = 4
AVN_MAX_STAGES = 6
AVN_MAX_WPARAMS = 8
AVN_MAX_GH = 32
AVN_TILE = 64
PHASES = ('warp', 'cov', 'factor', 'beta', 'unused4', 'alpha', 'kinv_grad', 'finalize', 'kxs', 'predict_var')

KERNEL_IDS = {'RBF': 0, 'Matern52': 1, 'Matern32': 2, 'Exponential': 3, 'RatQuad': 4}
OP_IDS = {'+': 0, '*': 1}


class WarpStage(C.Structure):
    _fields_ = [('op', C.c_int32), ('pidx', C.c_int32), ('c', C.c_double * 4)]


class WarpProg(C.Structure):
    _fields_ = [('nstages', C.c_int32), ('nparams', C.c_int32), ('st', WarpStage * AVN_MAX_STAGES)]


class ModelDesc(C.Structure):
    _fields_ = [('d', C.c_int32), ('nkern', C.c_int32), ('kern', C.c_int32 * AVN_MAX_KERN),
                ('op', C.c_int32 * AVN_MAX_KERN), ('noise', C.c_int32), ('jitter', C.c_double),
                ('xwarp', WarpProg * AVN_MAX_D), ('ywarp', WarpProg)]


class WsLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('npad', 'nb', 'xw', 'dxw', 'xs', 'x2', 'z', 'dz', 'wstat', 'kl', 't',
                                         'beta', 'alpha', 'gpart', 'gxpart', 'fpart', 'fflags', 'total')]


class Epilogue(C.Structure):
    _fields_ = [('mode', C.c_int32), ('deg', C.c_int32), ('normvar', C.c_int32), ('ei_max', C.c_int32),
                ('yopt', C.c_double), ('nodes', C.c_double * AVN_MAX_GH), ('weights', C.c_double * AVN_MAX_GH),
                ('yrev', WarpProg)]


# every symbol declared in include/avn_gp.h: name -> (restype, argtypes)
SYMBOLS = {
    'avn_last_error': (C.c_char_p, []),
    'avn_version': (C.c_int, []),
    'avn_gp_create': (C.c_int, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    'avn_gp_destroy': (None, [C.c_void_p]),
    'avn_gp_num_params': (C.c_int, [C.c_void_p]),
    'avn_gp_set_data': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    'avn_gp_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'avn_gp_workspace_layout': (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(WsLayout)]),
    'avn_gp_loglik_grad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    'avn_gp_host_staging_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'avn_gp_loglik_grad_host': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    'avn_gp_host_wait': (C.c_int, [C.c_void_p]),
    'avn_gp_set_streams': (C.c_int, [C.c_void_p, C.c_int]),
    'avn_gp_cov': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'avn_gp_state_bytes': (C.c_size_t, [C.c_void_p]),
    'avn_gp_factorize': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    'avn_gp_predict_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'avn_gp_predict': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Epilogue), C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'avn_gp_predict_grad_workspace_bytes': (C.c_size_t, [C.c_void_p, C.c_int64]),
    'avn_gp_predict_grad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Epilogue), C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    'avn_gp_append_workspace_bytes': (C.c_size_t, [C.c_void_p]),
    'avn_gp_append': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    'avn_gp_last_launch_count': (C.c_int64, [C.c_void_p]),
    'avn_gp_set_debug': (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    'avn_gp_set_profiling': (C.c_int, [C.c_void_p, C.c_int]),
    'avn_gp_phase_ms': (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the shared library once and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} not found: the CUDA extension is required (no CPU fallback exists). '
            'Build it with `python -c "import __graft_entry__ as g; g.build()"` from the repo root.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().avn_last_error().decode()


def make_prog(stages, nparams=None):
    """stages: list of (opcode, pidx, (c0..c3)) as produced by transform.wgp.program()."""
    p = WarpProg()
    stages = list(stages or [])
    if len(stages) > AVN_MAX_STAGES:
        raise ValueError(f'at most {AVN_MAX_STAGES} stages per composite warp')
    p.nstages = len(stages)
    npar = 0
    for i, (op, pidx, c) in enumerate(stages):
        p.st[i].op = int(op)
        p.st[i].pidx = int(pidx)
        for j in range(4):
            p.st[i].c[j] = float(c[j]) if j < len(c) else 0.0
    if nparams is None:
        from .transform import STAGES
        by_op = {v[0]: len(v[1]) for v in STAGES.values()}
        npar = sum(by_op.get(int(op), 0) for op, pidx, _ in stages if pidx >= 0)
    else:
        npar = int(nparams)
    if npar > AVN_MAX_WPARAMS:
        raise ValueError(f'at most {AVN_MAX_WPARAMS} learnable parameters per composite warp')
    p.nparams = npar
    return p

"""Device-side GP engine: thin Python object over the C ABI (``include/avn_gp.h``).

PyTorch is used for device memory, streams and (in ``dist.py``) the process group only; all
arithmetic happens in ``libavn_gp.so``.  One :class:`GPEngine` corresponds to one PyMC model of the
reference (``pm.Model()`` built in ``GPMCMC.__fit``, andvaranaut/gpmcmc.py:185-323): a kernel fold,
a noise flag, optional learnable input/output warps and the training data.
"""
import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from ._lib import KERNEL_IDS, OP_IDS


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class GPError(RuntimeError):
    pass


def _on_device(fn):
    """Run an engine method with the engine's device current (the calling thread's current device is restored
    afterwards): buffers, the stream handed to the library and the library's launches all belong to ``self.device``
    whatever thread or current device the call comes from (new host threads start on device 0)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


def check_info(info, what='avn_gp'):
    """info < 0 = the device aborted the factorisation (a dataflow wait timed out, include/avn_gp.h): an error, never
    data.  ``info`` is a host array or scalar."""
    if np.any(np.asarray(info) < 0):
        raise GPError(f'{what}: factorisation aborted on the device (a progress-flag wait of the factor kernel timed out; '
                      'the results are invalid) -- is another process time-slicing this GPU?')


class GPEngine:
    def __init__(self, nx, kerns=('RBF',), ops=(), noise=True, jitter=1e-6, xwarps=None, ywarp=None,
                 device=None):
        """xwarps: per input dimension ``None`` or a list of stage tuples (``wgp.program()``);
        ywarp: ``None`` or a list of stage tuples."""
        if not torch.cuda.is_available():
            raise GPError('GPEngine needs a CUDA device: there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        kerns = list(kerns)
        ops = list(ops)
        if len(ops) != len(kerns) - 1:
            raise ValueError('need one op between each pair of kernels')
        self.nx, self.kerns, self.ops, self.noise, self.jitter = nx, kerns, ops, bool(noise), float(jitter)
        desc = _lib.ModelDesc()
        desc.d = nx
        desc.nkern = len(kerns)
        for i, k in enumerate(kerns):
            desc.kern[i] = KERNEL_IDS[k]
        for i, o in enumerate(ops):
            desc.op[i] = OP_IDS[o]
        desc.noise = 1 if noise else 0
        desc.jitter = jitter
        self.n_iw = 0
        if xwarps is not None:
            if len(xwarps) != nx:
                raise ValueError('xwarps must have one entry per input dimension')
            for m, prog in enumerate(xwarps):
                if prog:
                    desc.xwarp[m] = _lib.make_prog(prog)
                    self.n_iw += desc.xwarp[m].nparams
        self.n_cw = 0
        if ywarp:
            desc.ywarp = _lib.make_prog(ywarp)
            self.n_cw = desc.ywarp.nparams
        self._desc = desc
        h = C.c_void_p()
        with torch.cuda.device(self.device):       # the handle is bound to the device current at create
            rc = self.lib.avn_gp_create(C.byref(desc), C.byref(h))
        if rc != 0:
            raise GPError(_lib.last_error())
        self._h = h
        self.P = self.lib.avn_gp_num_params(h)
        self.N = 0
        self._X = self._y = None
        self._ws = None
        self._state = None
        self._pws = None
        self.launches = 0

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            try:
                self.lib.avn_gp_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_debug(self, wait_bound_log2=26, fault=0):
        """wait bound of the factor kernel's flag waits (2^k polls) and fault injection (tests)."""
        if self.lib.avn_gp_set_debug(self._h, int(wait_bound_log2), int(fault)) != 0:
            raise GPError(_lib.last_error())

    def set_streams(self, max_groups):
        if self.lib.avn_gp_set_streams(self._h, int(max_groups)) != 0:
            raise GPError(_lib.last_error())

    def set_profiling(self, enable=True):
        if self.lib.avn_gp_set_profiling(self._h, 1 if enable else 0) != 0:
            raise GPError(_lib.last_error())

    @_on_device
    def phase_ms(self):
        """elapsed milliseconds per phase of the most recent call (synchronises on its last event)."""
        out = (C.c_double * len(_lib.PHASES))()
        if self.lib.avn_gp_phase_ms(self._h, out) != 0:
            raise GPError(_lib.last_error())
        return dict(zip(_lib.PHASES, list(out)))

    # -- layout helpers ----------------------------------------------------------------------
    def offsets(self):
        o, p = {}, 0
        if self.noise:
            o['gv'] = p
            p += 1
        o['l'] = p
        p += self.nx * len(self.kerns)
        o['kv'] = p
        p += len(self.kerns)
        o['iw'] = p
        p += self.n_iw
        o['cw'] = p
        p += self.n_cw
        if 'RatQuad' in self.kerns:
            o['alpha'] = p
            p += 1
        o['P'] = p
        return o

    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=torch.float64).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.device)

    @_on_device
    def set_data(self, X, y, _keep_state=False):
        X = self._dev(X)
        y = self._dev(y).reshape(-1)
        if X.ndim != 2 or X.shape[1] != self.nx or X.shape[0] != y.shape[0]:
            raise ValueError('X must be [N,nx] and y [N]')
        self._X, self._y, self.N = X, y, X.shape[0]
        rc = self.lib.avn_gp_set_data(self._h, _ptr(X), _ptr(y), self.N)
        if rc != 0:
            raise GPError(_lib.last_error())
        if not _keep_state:
            self._state = None

    @property
    def npad(self):
        return (self.N + _lib.AVN_TILE - 1) // _lib.AVN_TILE * _lib.AVN_TILE

    @_on_device
    def _workspace(self, B):
        need = self.lib.avn_gp_workspace_bytes(self._h, B)
        if need == 0:
            raise GPError('set_data first')
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def max_batch(self, budget_bytes):
        """largest B whose loglik workspace fits in ``budget_bytes``."""
        per = self.lib.avn_gp_workspace_bytes(self._h, 1)
        return max(1, int(budget_bytes // per))

    # -- hot path ----------------------------------------------------------------------------
    @_on_device
    def loglik_grad(self, theta, want_grad=True, out=None):
        """theta [B,P] (constrained).  Returns (ll [B], grad [B,P] or None, info [B]) device tensors."""
        theta = self._dev(theta)
        if theta.ndim == 1:
            theta = theta[None, :]
        B = theta.shape[0]
        if theta.shape[1] != self.P:
            raise ValueError(f'theta must have {self.P} columns')
        ws = self._workspace(B)
        if out is None:
            ll = torch.empty(B, dtype=torch.float64, device=self.device)
            grad = torch.empty(B, self.P, dtype=torch.float64, device=self.device) if want_grad else None
            info = torch.empty(B, dtype=torch.int32, device=self.device)
        else:
            ll, grad, info = out
        rc = self.lib.avn_gp_loglik_grad(self._h, _ptr(theta), B, _ptr(ll), _ptr(grad), _ptr(info), _ptr(ws),
                                         ws.numel(), self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        self._last_B = B
        return ll, grad, info

    def loglik_grad_host(self, theta, want_grad=True, _async=False):
        """host arrays in, host arrays out: theta [B,P] NumPy -> (ll [B], grad [B,P], info [B]) NumPy -- the call the
        optimiser / sampler drivers make once per step.  From the second consecutive call with the same batch size on
        it goes through ``avn_gp_loglik_grad_host``: the point is written into a pinned host buffer and the host->device
        copy, the launches of the evaluation and ONE packed device->host copy of ll, grad and info are a CUDA graph
        inside the library (captured once, replayed afterwards), so a step costs one C call.  A batch size seen for the
        first time (the drain of a sampler run, one-off calls) takes the same launches un-captured."""
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        B, P = theta.shape
        if P != self.P:
            raise ValueError(f'theta must have {self.P} columns')
        if getattr(self, '_host_last_B', None) != B:
            self._host_last_B = B
            return self._loglik_grad_host_plain(theta, want_grad)
        hc = getattr(self, '_hostcall', None)
        if hc is None or hc['B'] != B or hc['ws'] is not self._workspace(B):
            with torch.cuda.device(self.device):
                th_host = torch.empty(B, P, dtype=torch.float64).pin_memory()
                out_host = torch.empty(B * (P + 2), dtype=torch.float64).pin_memory()
                staging = torch.empty(self.lib.avn_gp_host_staging_bytes(self._h, B), dtype=torch.uint8, device=self.device)
                ws = self._workspace(B)
            h = out_host.numpy()
            hc = self._hostcall = dict(B=B, ws=ws, th_host=th_host, th_np=th_host.numpy(), out_host=out_host, staging=staging,
                                       ll=h[:B], grad=h[B:B + B * P].reshape(B, P), info=h[B + B * P:].view(np.int32)[:B],
                                       args=(self._h, _ptr(th_host), B, _ptr(out_host)),
                                       tail=(_ptr(staging), staging.numel(), _ptr(ws), ws.numel()))
        hc['th_np'][...] = theta
        hc['want_grad'] = want_grad
        rc = self.lib.avn_gp_loglik_grad_host(*hc['args'], (1 if want_grad else 0) | (2 if _async else 0), *hc['tail'],
                                              self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        self._last_B = B
        if _async:
            return hc
        return self._host_results(hc)

    @staticmethod
    def _host_results(hc):
        info = hc['info'].copy()
        check_info(info, 'avn_gp_loglik_grad')
        return hc['ll'].copy(), (hc['grad'].copy() if hc['want_grad'] else None), info

    def loglik_grad_host_begin(self, theta, want_grad=True):
        """asynchronous form of :meth:`loglik_grad_host` for the drivers: the evaluation is launched and the call returns a
        token; the caller computes its host-side terms (priors, Jacobians) while the device works and collects
        (ll, grad, info) with :meth:`loglik_grad_host_end`.  Batch sizes not yet captured are evaluated at once."""
        out = self.loglik_grad_host(theta, want_grad, _async=True)
        return out

    def loglik_grad_host_end(self, token):
        if isinstance(token, tuple):          # evaluated synchronously (first call of a batch size)
            return token
        if self.lib.avn_gp_host_wait(self._h) != 0:
            raise GPError(_lib.last_error())
        return self._host_results(token)

    @_on_device
    def _loglik_grad_host_plain(self, theta, want_grad=True):
        """the un-captured form of :meth:`loglik_grad_host`: pinned upload, ``avn_gp_loglik_grad`` on the current stream,
        one packed device->host copy, one stream synchronisation."""
        B, P = theta.shape
        n = B * (P + 2)
        if getattr(self, '_pack', None) is None or self._pack.numel() != n:
            self._pack = torch.empty(n, dtype=torch.float64, device=self.device)
            self._pack_host = torch.empty(n, dtype=torch.float64).pin_memory()
            self._theta_host = torch.empty(B, P, dtype=torch.float64).pin_memory()
        buf = self._pack
        out = (buf[:B], buf[B:B + B * P].view(B, P) if want_grad else None, buf[B + B * P:].view(torch.int32)[:B])
        self._theta_host.copy_(torch.from_numpy(theta))
        self.loglik_grad(self._theta_host.to(self.device, non_blocking=True), want_grad=want_grad, out=out)
        self._pack_host.copy_(buf, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        h = self._pack_host.numpy()
        info = h[B + B * P:].view(np.int32)[:B].copy()
        check_info(info, 'avn_gp_loglik_grad')
        return h[:B].copy(), (h[B:B + B * P].reshape(B, P).copy() if want_grad else None), info

    @_on_device
    def cov(self, theta):
        theta = self._dev(theta)
        if theta.ndim == 1:
            theta = theta[None, :]
        B = theta.shape[0]
        ws = self._workspace(B)
        K = torch.empty(B, self.npad, self.npad, dtype=torch.float64, device=self.device)
        rc = self.lib.avn_gp_cov(self._h, _ptr(theta), B, _ptr(K), _ptr(ws), ws.numel(), self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        return K

    @_on_device
    def debug_buffers(self, B=None):
        """views of the named workspace buffers of the last loglik_grad call (tests / profiling)."""
        B = B or self._last_B
        L = _lib.WsLayout()
        rc = self.lib.avn_gp_workspace_layout(self._h, B, C.byref(L))
        if rc != 0:
            raise GPError(_lib.last_error())
        npad, nb, d, nk = L.npad, L.nb, self.nx, len(self.kerns)

        def view(off, shape):
            n = int(np.prod(shape))
            return self._ws[off:off + n * 8].view(torch.float64).view(*shape)
        out = dict(npad=npad, nb=nb,
                   xw=view(L.xw, (B, npad, d)), xs=view(L.xs, (B, nk, npad, d)), x2=view(L.x2, (B, nk, npad)),
                   z=view(L.z, (B, npad)), wstat=view(L.wstat, (B, 16)), kl=view(L.kl, (B, npad, npad)),
                   t=view(L.t, (B, npad, npad)), beta=view(L.beta, (B, npad)), alpha=view(L.alpha, (B, npad)))
        if self.n_iw:
            out['dxw'] = view(L.dxw, (B, npad, d, 8))
            out['gxpart'] = view(L.gxpart, (B, nb, npad, d))
        if self.n_cw:
            out['dz'] = view(L.dz, (B, npad, 8))
        return out

    # -- predict -----------------------------------------------------------------------------
    @_on_device
    def factorize(self, theta):
        theta = self._dev(theta).reshape(-1)
        if theta.shape[0] != self.P:
            raise ValueError(f'theta must have {self.P} entries')
        ws = self._workspace(1)
        sb = self.lib.avn_gp_state_bytes(self._h)
        if self._state is None or self._state.numel() < sb:
            self._state = torch.empty(sb, dtype=torch.uint8, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        rc = self.lib.avn_gp_factorize(self._h, _ptr(theta), _ptr(self._state), self._state.numel(), _ptr(info),
                                       _ptr(ws), ws.numel(), self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        self._theta_fact = theta
        return info

    @_on_device
    def append(self, xnew, znew):
        """Extend the factorised state by one converted training point (hyperparameters unchanged): O(N^2) rank-1
        update of T = L^-1 and alpha in place (``avn_gp_append``); a full refactorisation only when the padded slab
        is full (every 64th point) .  Returns info (device int32 [1]): non-zero = new pivot not positive, in which
        case the engine is left on the OLD data set."""
        if self._state is None:
            raise GPError('factorize first')
        xnew = self._dev(xnew).reshape(-1)
        znew = self._dev(np.asarray(znew, dtype=np.float64).reshape(-1)[:1]) if not isinstance(znew, torch.Tensor) \
            else self._dev(znew).reshape(-1)[:1]
        if xnew.shape[0] != self.nx:
            raise ValueError('xnew must have nx entries')
        Xn = torch.cat([self._X, xnew[None, :]])
        yn = torch.cat([self._y, znew])
        need = self.lib.avn_gp_append_workspace_bytes(self._h)
        if self._pws is None or self._pws.numel() < need:
            self._pws = None
            self._pws = torch.empty(need, dtype=torch.uint8, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        rc = self.lib.avn_gp_append(self._h, _ptr(self._state), self._state.numel(), _ptr(xnew), _ptr(znew), _ptr(info),
                                    _ptr(self._pws), self._pws.numel(), self._stream())
        if rc < 0:
            raise GPError(_lib.last_error())
        if rc == 1:                       # slab full: npad grows, new layout
            Xo, yo = self._X, self._y
            self.set_data(Xn, yn)
            info = self.factorize(self._theta_fact)
            if int(info[0]) != 0:
                self.set_data(Xo, yo)
                self.factorize(self._theta_fact)
            return info
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        if int(info[0]) == 0:
            self.set_data(Xn, yn, _keep_state=True)
        return info

    @staticmethod
    def make_epilogue(mode='latent', deg=8, normvar=False, EIopt=None, yopt=0.0, yrev=None):
        e = _lib.Epilogue()
        e.mode = {'latent': 0, 'revert': 1, 'EI': 2}[mode]
        e.deg = deg
        e.normvar = 1 if normvar else 0
        e.ei_max = 1 if EIopt == 'max' else 0
        e.yopt = float(yopt)
        if e.mode != 0:
            if deg < 1 or deg > _lib.AVN_MAX_GH:
                raise ValueError(f'deg must be in [1,{_lib.AVN_MAX_GH}]')
            xi, wi = np.polynomial.hermite.hermgauss(deg)
            for i in range(deg):
                e.nodes[i] = xi[i]
                e.weights[i] = wi[i]
        if yrev:
            e.yrev = _lib.make_prog(yrev, nparams=0)
        return e

    @staticmethod
    def _one_tile_bytes(full, M):
        """smallest usable predict workspace: one 64-point column block (the library's size is linear in the panel width)"""
        cols = min((M + _lib.AVN_TILE - 1) // _lib.AVN_TILE * _lib.AVN_TILE, 148 * 64 * 2)
        return full // cols * _lib.AVN_TILE

    @_on_device
    def predict(self, Xs, epilogue=None, mean_add=None, max_ws_bytes=4 << 30):
        """Xs [M,nx] converted test points -> (mean [M], var [M]) device tensors."""
        if self._state is None:
            raise GPError('factorize first')
        Xs = self._dev(Xs)
        M = Xs.shape[0]
        if Xs.ndim != 2 or Xs.shape[1] != self.nx:
            raise ValueError('Xs must be [M,nx]')
        epi = epilogue if epilogue is not None else self.make_epilogue()
        full = self.lib.avn_gp_predict_workspace_bytes(self._h, M)
        need = max(min(full, max_ws_bytes), self._one_tile_bytes(full, M))
        if self._pws is None or self._pws.numel() < need:
            self._pws = None
            self._pws = torch.empty(need, dtype=torch.uint8, device=self.device)
        mean = torch.empty(M, dtype=torch.float64, device=self.device)
        var = torch.empty(M, dtype=torch.float64, device=self.device)
        madd = self._dev(mean_add).reshape(-1) if mean_add is not None else None
        rc = self.lib.avn_gp_predict(self._h, _ptr(self._state), _ptr(Xs), M, C.byref(epi), _ptr(madd), _ptr(mean),
                                     _ptr(var), _ptr(self._pws), self._pws.numel(), self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        return mean, var

    @_on_device
    def predict_grad(self, Xs, epilogue=None, mean_add=None, dmean_add=None, pred_noise=True, max_ws_bytes=4 << 30):
        """Xs [M,nx] converted query points -> (mean [M], var [M], dmean [M,nx], dvar [M,nx]) device tensors: the
        predictive graph of the BO refine step (gpmcmc.py:738-801) with its gradient w.r.t. the query points.
        ``pred_noise=False`` leaves ``gv`` out of the variance as that inline graph does."""
        if self._state is None:
            raise GPError('factorize first')
        Xs = self._dev(Xs)
        if Xs.ndim != 2 or Xs.shape[1] != self.nx:
            raise ValueError('Xs must be [M,nx]')
        M = Xs.shape[0]
        epi = epilogue if epilogue is not None else self.make_epilogue()
        full = self.lib.avn_gp_predict_grad_workspace_bytes(self._h, M)
        need = max(min(full, max_ws_bytes), self._one_tile_bytes(full, M))
        if self._pws is None or self._pws.numel() < need:
            self._pws = None
            self._pws = torch.empty(need, dtype=torch.uint8, device=self.device)
        mean = torch.empty(M, dtype=torch.float64, device=self.device)
        var = torch.empty(M, dtype=torch.float64, device=self.device)
        dmean = torch.empty(M, self.nx, dtype=torch.float64, device=self.device)
        dvar = torch.empty(M, self.nx, dtype=torch.float64, device=self.device)
        madd = self._dev(mean_add).reshape(-1) if mean_add is not None else None
        dmadd = self._dev(dmean_add).reshape(M, self.nx) if dmean_add is not None else None
        rc = self.lib.avn_gp_predict_grad(self._h, _ptr(self._state), _ptr(Xs), M, C.byref(epi), 1 if pred_noise else 0,
                                          _ptr(madd), _ptr(dmadd), _ptr(mean), _ptr(var), _ptr(dmean), _ptr(dvar),
                                          _ptr(self._pws), self._pws.numel(), self._stream())
        if rc != 0:
            raise GPError(_lib.last_error())
        self.launches = self.lib.avn_gp_last_launch_count(self._h)
        return mean, var, dmean, dvar

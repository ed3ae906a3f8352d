"""GPU tests of the failure paths and the device binding: an aborted factorisation is reported (info = -1), never
returned as data; an engine works from any host thread / current device; the predict cache of GPMCMC follows the
hyperparameter values."""
import os
import sys
import threading

import numpy as np
import pytest
import scipy.stats as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from andvaranaut_b200 import GPMCMC, maxmin, meanstd  # noqa: E402
from andvaranaut_b200.gp import GPEngine, GPError  # noqa: E402
from oracle.gp_oracle import ModelSpec  # noqa: E402
from cases import engine_args, synth  # noqa: E402


def _engine(N=300, device=None, seed=5):
    spec = ModelSpec(nx=3, kerns=['Matern52'], noise=True)
    X, y, th, Xs = synth(spec, N, seed, M=70)
    eng = GPEngine(**engine_args(spec), device=device)
    eng.set_data(X, y)
    return eng, th, Xs


def test_timed_out_wait_is_reported_not_returned():
    """fault injection: the panel tile P(0,0,1) of sample 0 is never published, so every task behind it waits until the
    (shortened) bound, raises the abort flag and the call reports info = -1 for all samples of the launch."""
    eng, th, Xs = _engine()
    thetas = np.stack([th, th * 1.01, th * 0.99])
    ll0, g0, i0 = (t.cpu().numpy() for t in eng.loglik_grad(thetas))
    assert np.all(i0 == 0)
    eng.set_debug(wait_bound_log2=12, fault=1)
    ll, g, info = (t.cpu().numpy() for t in eng.loglik_grad(thetas))
    assert np.all(info == -1) and np.all(np.isnan(ll)) and np.all(g == 0.0)
    with pytest.raises(GPError, match='aborted'):
        eng.loglik_grad_host(thetas)
    assert int(eng.factorize(th)[0]) == -1
    # the handle recovers: the flags are re-zeroed by every call
    eng.set_debug(wait_bound_log2=26, fault=0)
    ll1, g1, i1 = (t.cpu().numpy() for t in eng.loglik_grad(thetas))
    assert np.all(i1 == 0) and np.array_equal(ll1, ll0) and np.array_equal(g1, g0)
    assert int(eng.factorize(th)[0]) == 0


def test_engine_from_worker_threads():
    """drivers.find_map_multi evaluates from threading.Thread workers: same bits as the main thread."""
    eng, th, Xs = _engine()
    ref = eng.loglik_grad_host(th[None, :])
    out = {}

    def work(i):
        out[i] = eng.loglik_grad_host(th[None, :])
    for i in range(3):
        t = threading.Thread(target=work, args=(i,))
        t.start()
        t.join()
    for i in range(3):
        assert all(np.array_equal(a, b) for a, b in zip(out[i], ref))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_engine_on_second_device_with_another_current_device():
    """GPEngine(device='cuda:1') while cuda:0 is current, from the main thread and from a worker thread (which starts on
    device 0): every launch, buffer and stream belongs to cuda:1 and the results equal those of cuda:0 bit for bit."""
    torch.cuda.set_device(0)
    e0, th, Xs = _engine(device='cuda:0')
    e1, _, _ = _engine(device='cuda:1')
    assert torch.cuda.current_device() == 0
    r0 = e0.loglik_grad_host(th[None, :])
    r1 = e1.loglik_grad_host(th[None, :])
    assert torch.cuda.current_device() == 0
    assert all(np.array_equal(a, b) for a, b in zip(r0, r1))
    e0.factorize(th)
    e1.factorize(th)
    m0, v0 = e0.predict(Xs)
    m1, v1 = e1.predict(Xs)
    assert m1.device.index == 1 and torch.equal(m0.cpu(), m1.cpu()) and torch.equal(v0.cpu(), v1.cpu())
    out = []
    t = threading.Thread(target=lambda: out.append(e1.loglik_grad_host(th[None, :])))
    t.start()
    t.join()
    assert all(np.array_equal(a, b) for a, b in zip(out[0], r0))


def _fitted(tmp_path, n=80):
    space = [st.uniform(0, 2), st.uniform(1, 0.5)]

    def target(x):
        return np.array([x[0] ** 2 - x[0] - x[1] ** 2 * x[0] + x[1]])
    g = GPMCMC(kernel='Matern52', noise=True, nx=2, ny=1, priors=space, target=target, parallel=False, nproc=1,
               verbose=False, rundir=str(tmp_path / 'runs'))
    g.sample(n, seed=11)
    g.change_conrevs([maxmin(g.x[:, 0]), maxmin(g.x[:, 1])], [meanstd(g.y[:, 0])])
    g.fit()
    return g


def test_predict_cache_follows_hyperparameter_values(tmp_path):
    """the reference passes point=self.hypers to gp.predict on every call (gpmcmc.py:593-594): edited or loaded hypers
    take effect at once, also at unchanged data length, and in-place edits of the converted data are seen."""
    g = _fitted(tmp_path)
    xs = np.column_stack([np.linspace(0.1, 1.9, 33), np.linspace(1.05, 1.45, 33)])
    y0, v0 = g.predict(xs, return_var=True)
    eng0 = g._pred_cache['eng']
    y0b = g.predict(xs)
    assert g._pred_cache['eng'] is eng0 and np.array_equal(y0, y0b)      # unchanged state: cached factorisation
    h = {k: np.array(v, copy=True) for k, v in g.hypers.items()}
    h['l'] = h['l'] * 1.7
    h['l_log__'] = np.log(h['l'])
    g.hypers = h
    y1, v1 = g.predict(xs, return_var=True)
    assert g._pred_cache['eng'] is not eng0 and np.max(np.abs(y1 - y0)) > 1e-6
    fresh = _fitted(tmp_path)
    fresh.hypers = h
    y2, v2 = fresh.predict(xs, return_var=True)
    assert np.array_equal(y1, y2) and np.array_equal(v1, v2)
    eng1 = g._pred_cache['eng']
    g.yc[3, 0] += 0.25                                                    # in-place edit, same length
    y3 = g.predict(xs)
    assert g._pred_cache['eng'] is not eng1 and np.max(np.abs(y3 - y1)) > 1e-9

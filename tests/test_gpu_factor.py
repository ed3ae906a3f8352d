"""GPU tests of the persistent dataflow factorisation (csrc/factor.cuh): Cholesky L, T = L^-1, beta, log det
against LAPACK through the oracle, over the block-count / batch-size combinations that exercise the ticket
order, the progress flags and the ragged last block row.  Everything goes through the C ABI (GPEngine)."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import gp_oracle as go  # noqa: E402
import cases  # noqa: E402


def engine(spec):
    from andvaranaut_b200.gp import GPEngine
    return GPEngine(**cases.engine_args(spec))


# (N, B): one row, exactly one / two tiles, one row into a new tile, a last block with 1, 8, 9 and 63 valid rows,
# more block rows than resident samples, more samples than resident CTAs per step
@pytest.mark.parametrize('N,B', [(1, 3), (64, 1), (128, 2), (129, 1), (136, 5), (137, 3), (191, 2), (257, 37),
                                 (449, 1), (449, 9), (70, 600)])
def test_factor_against_lapack(N, B):
    spec = go.ModelSpec(nx=3, kerns=['Matern52'])
    X, y, th, _ = cases.synth(spec, max(N, 2), seed=100 + N)
    X, y = X[:N], y[:N]                  # (synth standardises y: a single point would be 0/0)
    rng = np.random.default_rng(N)
    thetas = th[None, :] * np.exp(0.05 * rng.normal(size=(B, len(th))))
    eng = engine(spec)
    eng.set_data(X, y)
    ll, grad, info = eng.loglik_grad(thetas)
    torch.cuda.synchronize()
    bufs = eng.debug_buffers()
    assert int(info.abs().sum()) == 0
    for b in sorted({0, B // 2, B - 1}):
        r = go.loglik(spec, thetas[b], X, y, keep=True)
        L = np.tril(bufs['kl'][b, :N, :N].cpu().numpy())
        Tm = np.tril(bufs['t'][b, :N, :N].cpu().numpy())
        Tref = sla.solve_triangular(r.L, np.eye(N), lower=True)
        assert np.max(np.abs(L - r.L)) <= 1e-12 * np.max(np.abs(r.L))
        assert np.max(np.abs(Tm - Tref)) <= 1e-10 * np.max(np.abs(Tref))
        assert np.max(np.abs(bufs['beta'][b, :N].cpu().numpy() - r.beta)) <= 1e-10 * np.max(np.abs(r.beta))
        assert abs(float(ll[b]) - r.ll) <= 1e-9 * abs(r.ll)
        gn = np.maximum(np.abs(r.grad), 1e-3 * np.max(np.abs(r.grad)))
        assert np.max(np.abs(grad[b].cpu().numpy() - r.grad) / gn) <= 1e-9
    # the padding of the last block row stays the identity in L and T (the kernels rely on it)
    npad = bufs['npad']
    if npad > N:
        one = torch.ones(npad - N, dtype=torch.float64, device=bufs['kl'].device)
        assert torch.equal(torch.diagonal(bufs['kl'][0])[N:], one) and torch.equal(torch.diagonal(bufs['t'][0])[N:], one)
        assert float(bufs['t'][0, N:, :N].abs().max()) == 0.0


def test_pivot_index_matches_lapack_and_neighbours_are_untouched():
    """indefinite matrices (negative "noise" pushed through the constrained-space ABI): info = LAPACK's order of
    the first non-positive leading minor, ll = -inf, grad = 0; every other sample of the batch is bit-identical
    to a run without the bad ones."""
    spec = go.ModelSpec(nx=2, kerns=['Matern32'], noise=True, jitter=0.0)
    N = 200
    X, y, th, _ = cases.synth(spec, N, seed=5)
    good = th[None, :] * np.exp(0.03 * np.random.default_rng(1).normal(size=(6, len(th))))
    bads = []
    for gv in (-0.3, -2e-3, -2e-4):
        t = th.copy()
        t[0] = gv
        bads.append(t)
    thetas = np.concatenate([good[:3], np.stack(bads), good[3:]])
    eng = engine(spec)
    eng.set_data(X, y)
    ll, grad, info = eng.loglik_grad(thetas)
    ll_ref, grad_ref, info_ref = eng.loglik_grad(good)
    info = info.cpu().numpy()
    orders = []
    for q, t in enumerate(bads):
        K = go.cov_matrix(spec, go.unpack(spec, t), X)
        K[np.diag_indices(N)] += t[0]
        order = int(sla.lapack.dpotrf(K, lower=1)[1])      # order of the first non-positive leading minor
        orders.append(order)
        assert order > 0 and info[3 + q] == order, (q, order, info)
        assert np.isneginf(float(ll[3 + q])) and float(grad[3 + q].abs().max()) == 0.0
    assert max(orders) > 64, orders                        # at least one failure beyond the first block column
    keep = [0, 1, 2, 6, 7, 8]
    assert np.all(info[keep] == 0) and int(info_ref.abs().sum()) == 0
    assert torch.equal(ll[keep], ll_ref) and torch.equal(grad[keep], grad_ref)


def test_repeated_runs_are_bit_identical_under_different_schedules():
    """the ticket scheduler hands tiles to whichever CTA is free; results must not depend on it."""
    spec = go.ModelSpec(nx=4, kerns=['Matern32'])
    X, y, th, _ = cases.synth(spec, 330, seed=9)
    thetas = th[None, :] * np.exp(0.05 * np.random.default_rng(3).normal(size=(150, len(th))))
    eng = engine(spec)
    eng.set_data(X, y)
    ll0, g0, _ = eng.loglik_grad(thetas)
    ll0, g0 = ll0.clone(), g0.clone()
    for groups in (1, 2, 4):
        eng.set_streams(groups)
        for _ in range(2):
            ll, g, info = eng.loglik_grad(thetas)
            assert torch.equal(ll, ll0) and torch.equal(g, g0) and int(info.abs().sum()) == 0
    # a sub-batch sees a different task numbering
    ll, g, _ = eng.loglik_grad(thetas[40:47])
    assert torch.equal(ll, ll0[40:47]) and torch.equal(g, g0[40:47])


def test_factorize_for_predict_matches_loglik_state():
    """avn_gp_factorize (B = 1, inverse kept in the state buffer) and the batched path share the kernel."""
    spec = go.ModelSpec(nx=5, kerns=['Matern52'])
    N = 700
    X, y, th, Xs = cases.synth(spec, N, seed=12, M=257)
    eng = engine(spec)
    eng.set_data(X, y)
    info = eng.factorize(th)
    assert int(info[0]) == 0
    mu, var = eng.predict(Xs)
    mu_r, var_r = go.predict(spec, th, X, y, Xs)
    kv = go.kdiag_total(spec, go.unpack(spec, th)['kv'])
    assert np.max(np.abs(mu.cpu().numpy() - mu_r)) <= 1e-8 * np.max(np.abs(mu_r))
    assert np.max(np.abs(var.cpu().numpy() - var_r) / np.maximum(np.abs(var_r), kv)) <= 1e-8


@pytest.mark.parametrize('N,B', [(40, 2), (64, 3), (100, 1), (130, 4), (449, 1), (1000, 2)])
def test_chain_launch_equals_throughput_launch(N, B, monkeypatch):
    """Few samples take the fused-panel instantiation of the factor kernel (one chain CTA per sample, Dpre tasks, parked
    S tiles); many samples the throughput instantiation.  Same arithmetic: forcing either mode on the same inputs
    (development knob AVN_FAC_FUSE, read at every launch) gives the same bits in L, T, ll and the gradient, for one
    block row, a ragged last block row and many block rows."""
    spec = go.ModelSpec(nx=3, kerns=['Matern52'])
    X, y, th, _ = cases.synth(spec, N, seed=300 + N)
    thetas = th[None, :] * np.exp(0.05 * np.random.default_rng(N).normal(size=(B, len(th))))
    eng = engine(spec)
    eng.set_data(X, y)
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('AVN_FAC_FUSE', mode)
        ll, grad, info = eng.loglik_grad(thetas)
        torch.cuda.synchronize()
        bufs = eng.debug_buffers()
        res[mode] = (ll.clone(), grad.clone(), info.clone(), bufs['kl'].clone(), bufs['t'].clone())
        assert int(info.abs().sum()) == 0
    npad = res['0'][3].shape[-1]
    tri = torch.tril(torch.ones(npad, npad, dtype=torch.bool, device=res['0'][3].device))
    assert torch.equal(res['0'][0], res['1'][0]) and torch.equal(res['0'][1], res['1'][1])
    assert torch.equal(res['0'][4][:, tri], res['1'][4][:, tri])          # T: lower triangle (the upper tiles are scratch)
    lt0, lt1 = res['0'][3][:, tri], res['1'][3][:, tri]
    assert torch.equal(lt0, lt1)

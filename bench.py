#!/usr/bin/env python
"""Benchmark of the GP inner loop (BASELINE.json metric: GP loglik+grad evals/s over batched hypers,
and predict pts/s) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Headline workload (config.workload): BASELINE.json configs[1] -- N=2000, d=8 ARD Matern-5/2 GP with
learnable input (uniform -> kumaraswamy per dimension) and output (log -> sal -> meanstd) warps, P = 30
hyperparameters; one "step" = one log-likelihood + gradient evaluation for a batch of B hyperparameter
samples per GPU (synthetic LHC data, seed 202).  Samples are independent units: with N GPUs
every rank evaluates its own B samples (weak scaling) and the per-shard likelihoods/gradients are
all-gathered over NCCL inside the timed region.

The JSON line also carries: `e2e` (same metric through GPEngine with HOST buffers, copies timed),
`roofline` (dominant kernel vs the FP64 tensor peak measured in this run), `cpu_baseline` (the NumPy/SciPy
oracle timed on the host cores), and `extra` (config 3 batched chains and config 4 predict pts/s).
`--impl reference` times the oracle port alone (the reference's own GP path needs PyMC, which cannot be
installed offline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


# ------------------------------------------------------------------------------------------------
# synthetic workloads
# ------------------------------------------------------------------------------------------------
def lhc(n, d, seed):
    from scipy.stats import qmc
    return qmc.LatinHypercube(d=d, seed=seed).random(n)


class Workload(tuple):
    """(engine kwargs, X, y, theta) of one synthetic configuration.  Only package code builds it: the GPU arm never
    touches ``oracle/``; ``oracle_spec(name)`` gives the CPU legs the oracle's description of the same model."""
    __slots__ = ()


def _warp_programs(d):
    import scipy.stats as st
    from andvaranaut_b200 import transform as T
    xw = [T.wgp(['uniform', 'kumaraswamy'], np.ones(2), y=np.linspace(0.1, 0.9, 8), xdist=st.uniform(0.0, 1.0)).program()
          for _ in range(d)]
    yw = T.wgp(['logarithm', 'sal', 'meanstd'], np.ones(4), y=np.linspace(0.5, 1.5, 8)).program()
    return xw, yw


def workload_c2(seed=202, N=2000, d=8):
    rng = np.random.default_rng(seed)
    X = lhc(N, d, seed)
    a = np.linspace(0.5, 2.0, d)
    y = np.exp(np.sum(np.sin(2 * np.pi * a * X), axis=1) / d + 0.5 * X[:, 0] * X[:, 1]) + 0.01 * rng.normal(size=N)
    xw, yw = _warp_programs(d)
    kw = dict(nx=d, kerns=['Matern52'], ops=[], noise=True, jitter=1e-6, xwarps=xw, ywarp=yw)
    # theta layout: [gv][l: d][kv][iwgp: 2 d][cwgp: 4]
    th = np.concatenate([[1e-4], 0.7 * np.ones(d), [1.5], np.ones(2 * d), [0.0, 1.0, 0.0, 1.0]])
    return Workload((kw, X, y, th))


def workload_c3(seed=303, N=1000, d=6):
    rng = np.random.default_rng(seed)
    X = lhc(N, d, seed)
    a = np.linspace(0.5, 2.0, d)
    y = np.sum(np.sin(2 * np.pi * a * X), axis=1) + 0.05 * rng.normal(size=N)
    y = (y - y.mean()) / y.std()
    kw = dict(nx=d, kerns=['RBF'], ops=[], noise=True, jitter=1e-6)
    th = np.concatenate([[2e-3], 0.6 * np.ones(d), [1.5]])
    return Workload((kw, X, y, th))


def workload_c4(seed=404, N=8192, d=10):
    rng = np.random.default_rng(seed)
    X = lhc(N, d, seed)
    y = np.sin(X @ np.linspace(0.5, 2.0, d)) + 0.01 * rng.normal(size=N)
    y = (y - y.mean()) / y.std()
    kw = dict(nx=d, kerns=['Matern52'], ops=[], noise=True, jitter=1e-6)
    th = np.concatenate([[1e-4], np.ones(d), [1.5]])
    return Workload((kw, X, y, th))


def oracle_spec(name):
    """the oracle's description of workload ``name`` (CPU baseline / reference legs only)."""
    from oracle.gp_oracle import ModelSpec
    if name == 'c2':
        return ModelSpec(nx=8, kerns=['Matern52'], noise=True, xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 8,
                         ywarp=['logarithm', 'sal', 'meanstd'])
    if name == 'c3':
        return ModelSpec(nx=6, kerns=['RBF'], noise=True)
    if name == 'c4':
        return ModelSpec(nx=10, kerns=['Matern52'], noise=True)
    raise ValueError(name)


def theta_cloud(th, B, seed, scale=0.1):
    rng = np.random.default_rng(seed)
    return th[None, :] * np.exp(scale * rng.normal(size=(B, len(th))))


def flops_ll(N, d):
    """algorithmic FP64 flops of one loglik+grad evaluation."""
    return N ** 3 + (3 * d + 9) * N ** 2


def flops_kinv_grad(N, d):
    """algorithmic flops of kinv_grad per sample: K^-1 tiles from T (N^3/3) + gradient contractions."""
    return N ** 3 / 3.0 + (3 * d + 4) * N ** 2 / 2.0


def flops_factor(N):
    """algorithmic flops of the dominant kernel per sample: Cholesky (N^3/3) + triangular inverse (N^3/3)."""
    return 2.0 * N ** 3 / 3.0


# dram__bytes_read.sum + dram__bytes_write.sum of ONE factor_kernel launch of the headline workload (B = 64) from the
# `ncu --set full` capture summarised in profiles/r01f_ncu_factor_kinv_b64.json
FACTOR_TRAFFIC_BYTES_B64 = 28.27e9 + 2.23e9


def flops_predict(N, d, deg=8):
    return N ** 2 + (3 * d + 14) * N + 40 * deg


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(sm)}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use all the host's cores."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count(), user_api='blas')
    except Exception:
        pass


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        info = [i for i in threadpool_info() if i.get('user_api') == 'blas']
        if info:
            return int(info[0]['num_threads']), info[0].get('internal_api', '?')
    except Exception:
        pass
    return os.cpu_count(), '?'


def cpu_baseline_ll(spec, X, y, thetas, budget_s=12.0, max_evals=64):
    """oracle loglik+grad timed on the host cores over a bounded sample of the same hyperparameter batch."""
    from oracle import gp_oracle as go
    go.loglik(spec, thetas[0], X, y)  # warm BLAS
    t0 = time.perf_counter()
    n = 0
    while n < min(len(thetas), max_evals):
        go.loglik(spec, thetas[n], X, y, want_grad=True)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import gp_oracle as go
    use_all_host_threads()
    _, X, y, th = workload_c2()
    spec = oracle_spec('c2')
    thetas = theta_cloud(th, max(args.steps + args.warmup, 4), seed=202)
    cores, api = blas_threads()
    per_step = 1  # one evaluation per step: ~1-2 s of multi-threaded LAPACK at N=2000
    for w in range(args.warmup):
        go.loglik(spec, thetas[w % len(thetas)], X, y)
    t0 = time.perf_counter()
    for s in range(args.steps):
        go.loglik(spec, thetas[(args.warmup + s) % len(thetas)], X, y, want_grad=True)
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    line = {
        'impl': 'reference', 'metric': 'gp_loglik_grad_evals_per_s', 'value': v, 'unit': 'evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'c2: N=2000 d=8 ARD Matern52 + learnable x/y warps, P=30', 'evals_per_step': per_step},
        'cpu_baseline': {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'blas': api,
                         'sample': f'{args.steps} sequential oracle loglik+grad evaluations of the c2 model '
                                   f'(NumPy/SciPy restatement of the PyMC path; PyMC itself is not installable offline)'},
        'e2e': {'value': v, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='hyperparameter samples per GPU per step (c2)')
    ap.add_argument('--streams', type=int, default=1, help='concurrent sample groups inside one call')
    ap.add_argument('--no-extra', action='store_true', help='skip the c3 / c4 side measurements')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from andvaranaut_b200.gp import GPEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- FP64 tensor peak of this GPU (cuBLAS DGEMM, plumbing only) -------------------------------
    def dgemm_peak(n=8192, reps=4):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12

    p64 = dgemm_peak()

    # ---- headline: c2 batched loglik+grad ----------------------------------------------------------
    kw2, X, y, th = workload_c2()
    N, d, P = X.shape[0], X.shape[1], len(th)
    B = args.batch
    eng = GPEngine(**kw2, device=dev)
    eng.set_data(X, y)
    eng.set_streams(args.streams)
    thetas = theta_cloud(th, B * world, seed=202)[rank * B:(rank + 1) * B]
    theta_dev = torch.as_tensor(thetas, device=dev)
    out = (torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, P, dtype=torch.float64, device=dev),
           torch.empty(B, dtype=torch.int32, device=dev))
    gathered = torch.empty(world * B, 1 + P, dtype=torch.float64, device=dev) if world > 1 else None
    packed = torch.empty(B, 1 + P, dtype=torch.float64, device=dev)

    def step():
        ll, grad, info = eng.loglik_grad(theta_dev, out=out)
        if world > 1:
            packed[:, 0] = ll
            packed[:, 1:] = grad
            dist.all_gather_into_tensor(gathered, packed)
        return ll

    for _ in range(args.warmup):
        step()
    launches_per_step = int(eng.launches)
    barrier()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    assert int(out[2].abs().sum()) == 0, 'non-PD sample in the benchmark batch'

    # ---- e2e: host buffers through the public engine call -------------------------------------------
    th_host = torch.as_tensor(thetas).pin_memory()
    ll_host = torch.empty(B, dtype=torch.float64).pin_memory()
    g_host = torch.empty(B, P, dtype=torch.float64).pin_memory()

    def step_e2e():
        ll, grad, info = eng.loglik_grad(th_host.to(dev, non_blocking=True), out=out)
        ll_host.copy_(ll, non_blocking=True)
        g_host.copy_(grad, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)

    # ---- per-kernel timing: the same step repeated args.steps times with CUDA events around every launch
    # (recorded by the library on the stream the kernels run on; one host sync per step to read them) ----
    eng.set_profiling(True)
    phases = {}
    reps = args.steps
    for _ in range(reps):
        eng.loglik_grad(theta_dev, out=out)
        pm = eng.phase_ms()
        for k, v in pm.items():
            phases[k] = phases.get(k, 0.0) + v / reps
    eng.set_profiling(False)
    fk_ms = phases['factor']
    fk_flops = B * flops_factor(N)
    achieved = fk_flops / (fk_ms * 1e-3) / 1e12
    kg_ach = B * flops_kinv_grad(N, d) / (phases['kinv_grad'] * 1e-3) / 1e12
    roofline = {
        'bound': 'tensor',
        'kernel': 'factor_kernel (persistent dataflow Cholesky + triangular inverse, FP64 DMMA tile products)',
        'achieved': achieved, 'peak': p64, 'unit': 'TFLOP/s', 'frac': achieved / p64,
        'traffic': FACTOR_TRAFFIC_BYTES_B64 if B == 64 else None,
        'peak_source': 'cuBLAS DGEMM fp64 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry; '
                       'DMMA issue peak measured 37.0 TF, profiles/r01_microbench_fp64.jsonl)',
        'algorithmic_flops_per_launch': fk_flops, 'kernel_ms': fk_ms,
        'second_kernel': {'kernel': 'kinv_grad_fast_kernel (K^-1 tiles fused with the gradient contraction)',
                          'achieved': kg_ach, 'frac': kg_ach / p64, 'kernel_ms': phases['kinv_grad'],
                          'algorithmic_flops_per_launch': B * flops_kinv_grad(N, d)},
        'step_frac': B * flops_ll(N, d) / (ms_per_step * 1e-3) / 1e12 / p64,
        'phase_ms': {k: round(v, 4) for k, v in phases.items() if v > 0},
    }

    line = {
        'metric': 'gp_loglik_grad_evals_per_s', 'value': value, 'unit': 'evals/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'c2: N={N} d={d} ARD Matern52 + learnable input (uniform,kumaraswamy) and output '
                               f'(log,sal,meanstd) warps, P={P}; B={B} hyperparameter samples per GPU per step',
                   'batch_per_gpu': B, 'parallelism': f'{world} x independent hyper batches, all_gather of [B,1+P]',
                   'l2': f'working set {eng._ws.numel() / 2**30:.1f} GiB per GPU, far larger than the 126 MB L2'},
        'clocks': clk.summary(),
        'e2e': {'value': e2e_value, 'unit': 'evals/s', 'h2d_bytes_per_step': int(B * P * 8),
                'd2h_bytes_per_step': int(B * (1 + P) * 8), 'ms_per_step': e2e_ms},
        'gpu_launches': launches_per_step * args.steps,
        'roofline': roofline,
    }

    # ---- side measurements: config 3 and config 4 ---------------------------------------------------
    extra = {}
    if not args.no_extra:
        # c3: 512 chains x N=1000, d=6, sharded over the ranks (strong scaling by definition of the config)
        kw3, X3, y3, th3 = workload_c3()
        B3 = 512 // world
        eng3 = GPEngine(**kw3, device=dev)
        eng3.set_data(X3, y3)
        eng3.set_streams(args.streams)
        t3 = torch.as_tensor(theta_cloud(th3, 512, seed=303)[rank * B3:(rank + 1) * B3], device=dev)
        for _ in range(3):
            o3 = eng3.loglik_grad(t3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k3 = max(3, args.steps // 2)
        e0.record()
        for _ in range(k3):
            o3 = eng3.loglik_grad(t3)
        e1.record()
        barrier()
        ms3 = max_over_ranks(e0.elapsed_time(e1)) / k3
        extra['c3_mcmc_chains'] = {'metric': 'gp_loglik_grad_evals_per_s', 'value': 512 / (ms3 * 1e-3), 'unit': 'evals/s',
                                   'ms_per_step': ms3, 'workload': '512 chains x N=1000 d=6 RBF, one ll+grad per chain',
                                   'scaling': 'strong', 'frac_of_fp64_peak': 512 * flops_ll(1000, 6) / (ms3 * 1e-3) / 1e12 / (p64 * world),
                                   'nonpd': int((o3[2] != 0).sum())}
        del eng3
        # c4: predict pts/s, N=8192 d=10, test blocks sharded (weak: M_sub points per GPU per step)
        kw4, X4, y4, th4 = workload_c4()
        eng4 = GPEngine(**kw4, device=dev)
        eng4.set_data(X4, y4)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng4.factorize(th4)
        e0.record()
        info4 = eng4.factorize(th4)
        e1.record()
        e1.synchronize()
        fact_ms = e0.elapsed_time(e1)
        Msub = 148 * 128 * 4
        gen = torch.Generator(device=dev)
        gen.manual_seed(405 + rank)
        Xs = torch.rand(Msub, 10, dtype=torch.float64, device=dev, generator=gen)
        epi = GPEngine.make_epilogue(mode='revert', deg=8, yrev=[(0, -1, (0.0, 1.0, 0.0, 0.0))])
        eng4.predict(Xs, epilogue=epi)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k4 = 2
        e0.record()
        for _ in range(k4):
            mu4, var4 = eng4.predict(Xs, epilogue=epi)
        e1.record()
        barrier()
        ms4 = max_over_ranks(e0.elapsed_time(e1)) / k4
        extra['c4_predict'] = {'metric': 'gp_predict_points_per_s', 'value': world * Msub / (ms4 * 1e-3), 'unit': 'pts/s',
                               'ms_per_step': ms4, 'points_per_gpu_per_step': Msub,
                               'workload': 'N=8192 d=10 Matern52, mean+variance+GH(8) reversion; 10M-point job streamed '
                                           f'in blocks of {Msub} per GPU', 'scaling': 'weak',
                               'frac_of_fp64_peak': world * Msub * flops_predict(8192, 10) / (ms4 * 1e-3) / 1e12 / (p64 * world),
                               'factorize_ms': fact_ms, 'info': int(info4[0])}
        if rank == 0 and world == 1 and not args.no_cpu:
            # CPU side of c4 (reported baseline): oracle predict with L factorised once (kinder than the reference,
            # which refactorises and recompiles on every call), 2 blocks of 4096 points, vectorised GH epilogue;
            # plus the reference's literal per-point GH loop (gpmcmc.py:549-563) on 10^4 points
            from oracle import gp_oracle as go
            use_all_host_threads()
            spec4 = oracle_spec('c4')
            cores, api = blas_threads()
            t0 = time.perf_counter()
            _, _, L4 = go.predict_blocked(spec4, th4, X4, y4, X4[:8], block=8)
            t_fact = time.perf_counter() - t0
            Xc = np.random.default_rng(405).uniform(size=(8192, 10))
            t0 = time.perf_counter()
            mu_c, var_c, _ = go.predict_blocked(spec4, th4, X4, y4, Xc, block=4096, L=L4)
            go.gh_stats(mu_c, var_c, lambda v: v, normvar=False)
            t_pred = time.perf_counter() - t0
            t0 = time.perf_counter()
            go.gh_stats_loop(np.resize(mu_c, 10000), np.resize(var_c, 10000), lambda v: v, normvar=False)
            t_loop = time.perf_counter() - t0
            extra['c4_predict']['cpu_baseline'] = {
                'value': 8192 / t_pred, 'unit': 'pts/s', 'cores': cores, 'kind': 'port', 'blas': api,
                'sample': '8192 of the test points in 2 blocks of 4096, L factorised once beforehand '
                          f'({t_fact:.1f} s, not counted), vectorised GH epilogue',
                'factorize_s': t_fact, 'reference_gh_loop_pts_per_s': 10000 / t_loop}
        # "next" rows of SURVEY 8f on the c4 / c3 data (rank 0): rank-1 append vs refactorisation, the inverse-problem
        # potential (value + gradient for a batch of candidate points), lock-step NUTS transitions
        if rank == 0:
            from andvaranaut_b200 import drivers
            from andvaranaut_b200.priors import ParamSpace
            from andvaranaut_b200.xpost import InverseLikelihood, XPosterior
            import scipy.stats as st
            nrow = {}
            Na = 8100
            enga = GPEngine(**kw4, device=dev)
            enga.set_data(X4[:Na], y4[:Na])
            enga.factorize(th4)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            enga.factorize(th4)
            torch.cuda.synchronize()
            t_fact = time.perf_counter() - t0
            xa, ya = torch.as_tensor(X4[Na:Na + 20], device=dev), torch.as_tensor(y4[Na:Na + 20], device=dev)
            enga.append(xa[0], ya[0:1])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(1, 20):
                info_a = enga.append(xa[i], ya[i:i + 1])
            torch.cuda.synchronize()
            t_app = (time.perf_counter() - t0) / 19
            nrow['f3_append'] = {'N': Na, 'append_ms': t_app * 1e3, 'refactorize_ms': t_fact * 1e3, 'info': int(info_a[0]),
                                 'algorithmic_bytes': 8 * Na * Na,
                                 'note': 'one new training point, hypers unchanged: avn_gp_append (4 launches, lower triangle '
                                         'of T read twice = 8 N^2 B) vs avn_gp_factorize of the enlarged set; host-timed incl. the info sync'}
            pot = InverseLikelihood(enga, lambda x: (x, np.ones_like(x)), np.array([0.3]), 1e-2, 0.0, 0.0)
            xpost = XPosterior([st.uniform(0, 1)] * 10, pot)
            zq = np.random.default_rng(7).normal(size=(256, 10))
            xpost.logp_dlogp(zq, True)
            t0 = time.perf_counter()
            for _ in range(3):
                xpost.logp_dlogp(zq, True)
            t_inv = (time.perf_counter() - t0) / 3
            nrow['f4_inverse'] = {'N': enga.N, 'candidates_per_call': 256, 'ms_per_call': t_inv * 1e3,
                                  'logp_dlogp_evals_per_s': 256 / t_inv,
                                  'note': 'inverse_opt potential + gradient w.r.t. x for 256 candidate points (restarts / '
                                          'chains) per call: one avn_gp_predict_grad + host Schur term, host-timed'}
            del enga, pot, xpost
            eng3 = GPEngine(**kw3, device=dev)
            eng3.set_data(X3, y3)
            sp3 = ParamSpace(6, 1, True)
            post3 = drivers.Posterior(eng3, sp3)
            z0 = sp3.z_from_theta(th3)
            t0 = time.perf_counter()
            tr = drivers.sample(post3, draws=4, tune=4, chains=128, seed=1, start_z=z0[None, :], init_jitter=0.05,
                                max_treedepth=4)
            t_nuts = time.perf_counter() - t0
            nrow['f1_nuts'] = {'chains': 128, 'N': 1000, 'transitions': 8, 'max_treedepth': 4, 'seconds': t_nuts,
                               'leapfrog_evals': int(post3.n_eval), 'device_calls': int(post3.n_calls),
                               'evals_per_s': post3.n_eval / t_nuts,
                               'mean_tree_steps': float(tr.sample_stats['n_steps'].mean()),
                               'note': 'lock-step NUTS through drivers.sample on the c3 model: every leapfrog of the '
                                       'still-growing chains is one batched avn_gp_loglik_grad call; host-timed end to end'}
            del eng3
            extra['next_rows'] = nrow
        # c1 (the reference's own tutorial case) and c5 through the GPMCMC API (rank 0 only: sequential optimisers)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import c1_tutorial_probe
            extra['c1_tutorial'] = c1_tutorial_probe.run()
            import c5_bo_probe
            extra['c5_bo'] = {'metric': 'bo_iterations_per_s', 'workload': 'd=12, 4096 LHC candidates -> EI -> argmax -> '
                              'append -> warm-started MAP refit, at three training-set sizes (3 iterations each)',
                              'sizes': [c5_bo_probe.run(n) for n in (256, 1024, 4096)]}
        line['extra'] = extra

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import gp_oracle as go
        use_all_host_threads()
        spec = oracle_spec('c2')
        r0 = go.loglik(spec, th, X, y, want_grad=False, keep=True)
        sv = np.linalg.svd(r0.L, compute_uv=False)
        line['config']['cond_K_nominal_theta'] = float((sv[0] / sv[-1]) ** 2)
        cores, api = blas_threads()
        v, n, dt = cpu_baseline_ll(spec, X, y, thetas)
        line['cpu_baseline'] = {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'blas': api,
                                'sample': f'{n} of the {B} hyperparameter samples of one step, sequential, all BLAS threads '
                                          f'({dt:.1f} s); oracle = NumPy/SciPy restatement of the PyMC path'}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

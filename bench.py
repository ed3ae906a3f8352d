#!/usr/bin/env python
"""Benchmark of the GP inner loop (BASELINE.json metric: GP loglik+grad evals/s over batched hypers,
and predict pts/s) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Headline workload (config.workload): BASELINE.json configs[1] -- N=2000, d=8 ARD Matern-5/2 GP with
learnable input (uniform -> kumaraswamy per dimension) and output (log -> sal -> meanstd) warps, P = 30
hyperparameters; one "step" = one log-likelihood + gradient evaluation for a batch of B hyperparameter
samples per GPU (synthetic LHC data, seed 202).  Samples are independent units: with N GPUs
every rank evaluates its own B samples (weak scaling) through ``Shard.loglik_grad_dev`` and the per-shard
likelihoods / gradients are all-gathered over NCCL inside the timed region.

Everything the driver keeps lives under its known keys:
  `e2e`            the same metric through the host-array API (``GPEngine.loglik_grad_host`` / ``Shard.loglik_grad``),
                   copies timed
  `roofline`       dominant kernel vs the FP64 DGEMM rate measured in this run, and `roofline.per_config`:
                   one self-contained record per BASELINE config (c1, c2 as the reference runs it, c3, c4, c5) and
                   for the "next" rows -- each with value / unit / ms / frac / scaling and, where they apply, its
                   own `e2e` and `cpu_baseline`
  `cpu_baseline`   the NumPy/SciPy oracle timed on the host cores (rank 0, N = 1 only)
`--impl reference` times the oracle port alone (the reference's own GP path needs PyMC, which cannot be
installed offline).  Both arms print the SAME `config`.
"""
import argparse
import glob
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
# synthetic workloads -- shared with tests/ so that the timed inputs are the verified inputs
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
    """returns the Workload and the (a, b) of the output conversion z = a + b y (meanstd) for the GH reversion."""
    rng = np.random.default_rng(seed)
    X = lhc(N, d, seed)
    yraw = 3.0 + 2.0 * np.sin(X @ np.linspace(0.5, 2.0, d)) + 0.02 * rng.normal(size=N)
    m, s = np.mean(yraw), np.std(yraw)
    ab = (-m / s, 1 / s)
    y = ab[0] + ab[1] * yraw
    kw = dict(nx=d, kerns=['Matern52'], ops=[], noise=True, jitter=1e-6)
    th = np.concatenate([[1e-4], np.ones(d), [1.5]])
    return Workload((kw, X, y, th)), ab


def c4_test_points(M, seed=405, d=10):
    return np.random.default_rng(seed).uniform(size=(M, d))


def workload_c5(N=4096, d=12, seed=505):
    """BO acquisition at training-set size N: LHC data of the C5 target, meanstd output conversion, plausible fitted
    hyperparameters; returns the Workload, (a, b) of the conversion and yopt (raw minimum, opt_type='min')."""
    X = lhc(N, d, seed)
    yraw = np.sum((X - 0.3) ** 2, axis=1) + np.sin(5.0 * X[:, 0])
    m, s = np.mean(yraw), np.std(yraw)
    ab = (-m / s, 1 / s)
    y = ab[0] + ab[1] * yraw
    kw = dict(nx=d, kerns=['Matern52'], ops=[], noise=True, jitter=1e-6)
    th = np.concatenate([[1e-4], np.r_[0.8, 3.0 * np.ones(d - 1)], [4.0]])
    return Workload((kw, X, y, th)), ab, float(yraw.min())


def c5_candidates(M=4096, d=12, seed=506):
    return lhc(M, d, seed)


def oracle_spec(name):
    """the oracle's description of workload ``name`` (CPU baseline / reference legs and tests only)."""
    from oracle.gp_oracle import ModelSpec
    if name == 'c2':
        return ModelSpec(nx=8, kerns=['Matern52'], noise=True, xwarps=[(['uniform', 'kumaraswamy'], (0.0, 1.0))] * 8,
                         ywarp=['logarithm', 'sal', 'meanstd'])
    if name == 'c3':
        return ModelSpec(nx=6, kerns=['RBF'], noise=True)
    if name == 'c4':
        return ModelSpec(nx=10, kerns=['Matern52'], noise=True)
    if name == 'c5':
        return ModelSpec(nx=12, kerns=['Matern52'], noise=True)
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


def flops_predict(N, d, deg=8):
    return N ** 2 + (3 * d + 14) * N + 40 * deg


def bytes_cov(N):
    """algorithmic HBM bytes of the covariance build per sample: the lower block triangle of K written once."""
    npad = (N + 63) // 64 * 64
    nb = npad // 64
    return nb * (nb + 1) // 2 * 64 * 64 * 8


HEADLINE = 'c2: N=2000 d=8 ARD Matern52 + learnable input (uniform,kumaraswamy) and output (log,sal,meanstd) warps, P=30'


def headline_config(B, world):
    """identical in both arms (the driver compares them)."""
    return {'workload': HEADLINE, 'batch_per_gpu': B,
            'parallelism': f'{world} x independent hyperparameter batches (weak scaling), all_gather of [B,2+P]',
            'l2': 'inputs larger than L2: the working set of one step is 4.4 GiB per GPU at B=64 (126 MB L2)'}


def traffic_from_profiles(kernel='factor_kernel', pattern='r*_ncu_*factor*_b64*.json'):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of ``kernel`` from the NEWEST committed `ncu --set full`
    summary of the headline workload (profiles/, written by tools/ncu_summary.py); (None, None) if there is none."""
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', pattern)), reverse=True):
        try:
            for rec in json.load(open(path)):
                if rec.get('kernel', '').startswith(kernel) or f' {kernel}' in rec.get('kernel', ''):
                    tot = rec['dram_read'] * unit[rec.get('dram_read_unit', 'byte')] + \
                        rec['dram_write'] * unit[rec.get('dram_write_unit', 'byte')]
                    return tot, os.path.relpath(path, ROOT)
        except (OSError, ValueError, KeyError):
            continue
    return None, None


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


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


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
    world = int(os.environ.get('WORLD_SIZE', '1'))
    thetas = theta_cloud(th, args.batch * world, seed=202)      # the batch our arm evaluates; a step = one of its rows
    cores, api = blas_threads()
    for w in range(args.warmup):
        go.loglik(spec, thetas[w % len(thetas)], X, y)
    t0 = time.perf_counter()
    for s in range(args.steps):
        go.loglik(spec, thetas[(args.warmup + s) % len(thetas)], X, y, want_grad=True)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'gp_loglik_grad_evals_per_s', 'value': v, 'unit': 'evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': headline_config(args.batch, world),
        'cpu_baseline': {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'blas': api,
                         'sample': f'each step = ONE row of the batch (sequential oracle loglik+grad evaluations, all BLAS '
                                   f'threads); {args.steps} rows timed.  Oracle = NumPy/SciPy restatement of the PyMC '
                                   f'path; PyMC itself is not installable offline'},
        'e2e': {'value': v, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


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
    ap.add_argument('--no-extra', action='store_true', help='headline only: skip the per-config records')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baselines')
    ap.add_argument('--c4-points', type=int, default=1 << 20, help='test points per GPU of the c4 host-to-host job')
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: everything else written to fd 1 during the run (NCCL's version banner,
    # library chatter) goes to stderr, the JSON line is written to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from andvaranaut_b200.dist import Shard
    from andvaranaut_b200.gp import GPEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    shard = Shard()
    cpu_legs = rank == 0 and world == 1 and not args.no_cpu

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

    def timed(fn, reps, warm=1):
        """device time of `reps` calls of fn (CUDA events on the current stream, barrier + sync both sides), max over
        ranks, per call, in ms"""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps

    def timed_host(fn, reps, warm=1):
        """wall time of host-to-host calls (every call ends with its results on the host), max over ranks, per call, ms"""
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) * 1e3) / reps

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
    peak_source = ('in-run cuBLAS DGEMM fp64 8192^3 (MEASURED_PEAKS.json has no FP64 entry and the profiling guide no FP64 '
                   'fallback; DMMA issue peak measured 37.0 TF, profiles/r01_microbench_fp64.jsonl)')
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        hbm_source = 'MEASURED_PEAKS.json hbm_gbs'
    except (OSError, ValueError, KeyError):
        hbm_peak, hbm_source = 6500.0, 'fallback (B200_PROFILING.md: measured copy bandwidth of this pool)'

    # ---- headline: c2 batched loglik+grad ----------------------------------------------------------
    kw2, X, y, th = workload_c2()
    N, d, P = X.shape[0], X.shape[1], len(th)
    B = args.batch
    eng = GPEngine(**kw2, device=dev)
    eng.set_data(X, y)
    eng.set_streams(args.streams)
    thetas_all = theta_cloud(th, B * world, seed=202)
    thetas = thetas_all[rank * B:(rank + 1) * B]
    theta_dev = torch.as_tensor(thetas, device=dev)
    packed = torch.zeros(B, P + 2, dtype=torch.float64, device=dev)
    gathered = torch.empty(world * B, P + 2, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        return shard.loglik_grad_dev(eng, theta_dev, B * world, packed=packed, gathered=gathered)

    for _ in range(args.warmup):
        res = step()
    launches_per_step = int(eng.launches)
    barrier()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = step()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    assert int(res[:, 1 + P].abs().sum()) == 0, 'non-PD or aborted sample in the benchmark batch'

    # ---- e2e: host arrays through the public call ---------------------------------------------------
    if world == 1:
        def step_e2e():
            return eng.loglik_grad_host(thetas)
        d2h = B * (P + 2) * 8
    else:
        def step_e2e():
            return shard.loglik_grad(eng, thetas_all)
        d2h = world * B * (P + 2) * 8
    e2e_ms = timed_host(step_e2e, args.steps, warm=3)   # warm-up covers the one-off capture of the host call as a CUDA graph
    e2e_value = world * B / (e2e_ms * 1e-3)

    # ---- per-kernel timing: the same step repeated with CUDA events around every launch (recorded by the
    # library on the stream the kernels run on; one host sync per step to read them) ----
    eng.set_profiling(True)
    phases = {}
    reps = args.steps
    with ClockSampler(local) as clk_ph:
        for _ in range(reps):
            eng.loglik_grad(theta_dev)      # the recorded call runs right behind another one, as inside the timed region
            eng.loglik_grad(theta_dev)      # (the host sync that reads the events would otherwise let the GPU idle first)
            for k, v in eng.phase_ms().items():
                phases[k] = phases.get(k, 0.0) + v / reps
    eng.set_profiling(False)
    fk_ms = phases['factor']
    fk_flops = B * flops_factor(N)
    achieved = fk_flops / (fk_ms * 1e-3) / 1e12
    kg_ach = B * flops_kinv_grad(N, d) / (phases['kinv_grad'] * 1e-3) / 1e12
    cov_gbs = B * bytes_cov(N) / (phases['cov'] * 1e-3) / 1e9
    traffic, traffic_src = traffic_from_profiles() if B == 64 else (None, None)
    roofline = {
        'bound': 'tensor',
        'kernel': 'factor_kernel (persistent dataflow Cholesky + triangular inverse, FP64 DMMA tile products)',
        'achieved': achieved, 'peak': p64, 'unit': 'TFLOP/s', 'frac': achieved / p64,
        'traffic': traffic, 'traffic_source': traffic_src,
        'peak_source': peak_source,
        'algorithmic_flops_per_launch': fk_flops, 'kernel_ms': fk_ms,
        'second_kernel': {'kernel': 'kinv_grad_fast_kernel (K^-1 tiles fused with the gradient contraction)',
                          'achieved': kg_ach, 'frac': kg_ach / p64, 'kernel_ms': phases['kinv_grad'],
                          'algorithmic_flops_per_launch': B * flops_kinv_grad(N, d)},
        'cov_kernel': {'kernel': 'cov1_kernel<Matern52> (covariance build, lower block triangle written once)',
                       'bound': 'hbm', 'achieved': cov_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': cov_gbs / hbm_peak,
                       'peak_source': hbm_source, 'kernel_ms': phases['cov'],
                       'algorithmic_bytes_per_launch': B * bytes_cov(N)},
        'step_frac': B * flops_ll(N, d) / (ms_per_step * 1e-3) / 1e12 / p64,
        'phase_ms': {k: round(v, 4) for k, v in phases.items() if v > 0},
        'phase_clocks': clk_ph.summary(),
    }

    line = {
        'metric': 'gp_loglik_grad_evals_per_s', 'value': value, 'unit': 'evals/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': headline_config(B, world),
        'clocks': clk.summary(),
        'e2e': {'value': e2e_value, 'unit': 'evals/s', 'h2d_bytes_per_step': int(B * P * 8),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms,
                'api': 'GPEngine.loglik_grad_host (pinned, one packed copy each way)' if world == 1 else
                       'Shard.loglik_grad (NumPy in, all_gather over NCCL, NumPy out on every rank)'},
        'gpu_launches': launches_per_step * args.steps,
        'roofline': roofline,
    }
    del eng, packed, gathered
    torch.cuda.empty_cache()

    per = {}
    if not args.no_extra:
        per_config(per, args, torch, dist, shard, GPEngine, dev, rank, world, p64, hbm_peak, timed, timed_host, barrier,
                   cpu_legs)
        roofline['per_config'] = per

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------
    if cpu_legs:
        from oracle import gp_oracle as go
        use_all_host_threads()
        spec = oracle_spec('c2')
        r0 = go.loglik(spec, th, X, y, want_grad=False, keep=True)
        sv = np.linalg.svd(r0.L, compute_uv=False)
        cores, api = blas_threads()
        v, n, dt = cpu_baseline_ll(spec, X, y, thetas)
        line['cpu_baseline'] = {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'blas': api,
                                'sample': f'{n} of the {B} hyperparameter samples of one step, sequential, all BLAS threads '
                                          f'({dt:.1f} s); oracle = NumPy/SciPy restatement of the PyMC path',
                                'cond_K_nominal_theta': float((sv[0] / sv[-1]) ** 2)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def per_config(per, args, torch, dist, shard, GPEngine, dev, rank, world, p64, hbm_peak, timed, timed_host, barrier, cpu_legs):
    """one self-contained record per BASELINE config and per "next" row, all through the product API:
    Shard.loglik_grad(_dev) / Shard.predict(_dev) / drivers.sample(shard=...) / GPMCMC; multi-GPU legs carry their
    gather inside the timed region."""
    from andvaranaut_b200 import drivers
    from andvaranaut_b200.priors import ParamSpace
    steps = max(3, args.steps // 2)

    # ---- c3: 512 chains x N=1000, d=6, sharded over the ranks (STRONG scaling by definition of the config) ----
    kw3, X3, y3, th3 = workload_c3()
    C3 = 512
    eng3 = GPEngine(**kw3, device=dev)
    eng3.set_data(X3, y3)
    t3_all = theta_cloud(th3, C3, seed=303)
    lo, hi, per3 = shard.bounds(C3)
    t3 = torch.as_tensor(t3_all[lo:hi], device=dev)
    P3 = t3_all.shape[1]
    pk = torch.zeros(per3, P3 + 2, dtype=torch.float64, device=dev)
    ga = torch.empty(world * per3, P3 + 2, dtype=torch.float64, device=dev) if world > 1 else None
    out3 = []

    def step3():
        out3[:] = [shard.loglik_grad_dev(eng3, t3, C3, packed=pk, gathered=ga)]
    ms3 = timed(step3, steps, warm=3)
    launches3 = int(eng3.launches)
    nonpd = int((out3[0][:, 1 + P3] != 0).sum())
    ms3_e2e = timed_host(lambda: shard.loglik_grad(eng3, t3_all) if world > 1 else eng3.loglik_grad_host(t3_all), steps, warm=3)
    eng3.set_profiling(True)
    eng3.loglik_grad(t3)
    ph3 = eng3.phase_ms()
    eng3.set_profiling(False)
    rec3 = {'metric': 'gp_loglik_grad_evals_per_s', 'value': C3 / (ms3 * 1e-3), 'unit': 'evals/s', 'ms_per_step': ms3,
            'workload': 'c3: 512 chains x N=1000 d=6 RBF + noise, one ll+grad per chain per step',
            'scaling': 'strong', 'n_gpus': world, 'chains_per_gpu': hi - lo, 'bound': 'tensor',
            'frac': C3 * flops_ll(1000, 6) / (ms3 * 1e-3) / 1e12 / (p64 * world),
            'factor_kernel_frac': (hi - lo) * flops_factor(1000) / (ph3['factor'] * 1e-3) / 1e12 / p64,
            'phase_ms_rank0': {k: round(v, 4) for k, v in ph3.items() if v > 0},
            'api': 'Shard.loglik_grad_dev: per-rank avn_gp_loglik_grad + all_gather_into_tensor of [512,2+P] inside the timed region',
            'gpu_launches_per_step': launches3, 'nonpd': nonpd,
            'e2e': {'value': C3 / (ms3_e2e * 1e-3), 'unit': 'evals/s', 'ms_per_step': ms3_e2e,
                    'h2d_bytes_per_step': int((hi - lo) * P3 * 8), 'd2h_bytes_per_step': int(C3 * (P3 + 2) * 8),
                    'api': 'Shard.loglik_grad (NumPy in / out on every rank)' if world > 1 else 'GPEngine.loglik_grad_host'},
            'limiter': 'no collective cost (127 KB gathered per step); per-GPU efficiency falls with fewer chains per GPU: '
                       'fewer tile tasks per block step of the persistent factor kernel (see DESIGN 5)'}
    if cpu_legs:
        use_all_host_threads()
        cores, api = blas_threads()
        v, n, dt = cpu_baseline_ll(oracle_spec('c3'), X3, y3, t3_all, budget_s=5.0, max_evals=48)
        rec3['cpu_baseline'] = {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'blas': api,
                                'sample': f'{n} of the 512 chains, sequential ({dt:.1f} s)'}
    per['c3'] = rec3

    # f1: NUTS with continuous batching over sharded chains through drivers.sample (every rank runs the same host sampler on the same
    # seed; each leapfrog round is one Shard.loglik_grad = per-rank device call + all_gather)
    sp3 = ParamSpace(6, 1, True)
    post3 = drivers.Posterior(eng3, sp3, shard=shard if world > 1 else None)
    z0 = sp3.z_from_theta(th3)
    barrier()
    t0 = time.perf_counter()
    tr = drivers.sample(post3, draws=16, tune=16, chains=128, seed=1, start_z=z0[None, :], init_jitter=0.05, max_treedepth=4)
    barrier()
    t_nuts = time.perf_counter() - t0
    per['f1_nuts'] = {'metric': 'nuts_leapfrog_evals_per_s', 'value': post3.n_eval / t_nuts, 'unit': 'evals/s',
                      'workload': 'c3 model: 128 chains, 32 NUTS transitions (16 tuning + 16 draws), max_treedepth 4, through '
                                  'drivers.sample (continuous batching; the drain at the end of the run is inside the numbers)',
                      'scaling': 'strong', 'n_gpus': world, 'seconds': t_nuts, 'leapfrog_evals': int(post3.n_eval),
                      'device_calls': int(post3.n_calls), 'mean_batch_per_call': post3.n_eval / max(post3.n_calls, 1),
                      'active_fraction': post3.n_eval / max(post3.n_calls * 128, 1),
                      'mean_tree_steps': float(tr.sample_stats['n_steps'].mean()),
                      'api': 'drivers.sample(Posterior(engine, space, shard=Shard()))' if world > 1 else 'drivers.sample'}
    del eng3, post3, pk, ga
    torch.cuda.empty_cache()

    # ---- c4: predict pts/s, N=8192 d=10, test blocks sharded (WEAK: M points per GPU) --------------------------
    (kw4, X4, y4, th4), ab4 = workload_c4()
    from andvaranaut_b200 import transform as T
    yrev4 = [(T.OP_AFFINE_CONST, -1, (ab4[0], ab4[1], 0.0, 0.0))]
    eng4 = GPEngine(**kw4, device=dev)
    eng4.set_data(X4, y4)
    eng4.factorize(th4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    info4 = eng4.factorize(th4)
    e1.record()
    e1.synchronize()
    fact_ms = e0.elapsed_time(e1)
    Msub = 148 * 128 * 4
    gen = torch.Generator(device=dev)
    gen.manual_seed(405 + rank)
    Xs = torch.rand(Msub, 10, dtype=torch.float64, device=dev, generator=gen)
    epi = GPEngine.make_epilogue(mode='revert', deg=8, yrev=yrev4)
    g4 = torch.empty(world * Msub, 2, dtype=torch.float64, device=dev) if world > 1 else None
    ms4 = timed(lambda: shard.predict_dev(eng4, Xs, world * Msub, gathered=g4, epilogue=epi), 2, warm=1)
    launches4 = int(eng4.launches)
    eng4.set_profiling(True)
    eng4.predict(Xs, epilogue=epi)
    ph4 = eng4.phase_ms()
    eng4.set_profiling(False)
    npad4 = eng4.npad
    # the launches of one call are panels of <= 18944 points; phase_ms holds the LAST panel of the call
    last_cols = Msub - (Msub - 1) // 18944 * 18944
    kxs_gbs = npad4 * last_cols * 8 / (ph4['kxs'] * 1e-3) / 1e9
    rec4 = {'metric': 'gp_predict_points_per_s', 'value': world * Msub / (ms4 * 1e-3), 'unit': 'pts/s', 'ms_per_step': ms4,
            'workload': 'c4: N=8192 d=10 Matern52 + noise, mean + variance + GH(8) reversion (meanstd); the 10M-point job '
                        f'is streamed in blocks of {Msub} points per GPU per step',
            'scaling': 'weak', 'n_gpus': world, 'points_per_gpu_per_step': Msub, 'bound': 'tensor',
            'kernel': 'predict_var_kernel (V = T K_xs by DMMA, column sums of V^2, GH epilogue)',
            'frac': world * Msub * flops_predict(8192, 10) / (ms4 * 1e-3) / 1e12 / (p64 * world),
            'kxs_kernel': {'bound': 'fp64 pipe / hbm write', 'achieved': kxs_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                           'frac': kxs_gbs / hbm_peak, 'kernel_ms': ph4['kxs'],
                           'algorithmic_bytes_per_launch': npad4 * last_cols * 8,
                           'note': 'K_xs panel written once; (3d+10) FP64 flops per 8 bytes: FP64-pipe bound, not HBM'},
            'phase_ms_last_panel': {k: round(v, 4) for k, v in ph4.items() if v > 0},
            'api': 'Shard.predict_dev: per-rank avn_gp_predict + all_gather_into_tensor of [M,2] inside the timed region',
            'gpu_launches_per_step': launches4, 'factorize_ms': fact_ms, 'info': int(info4[0])}
    # the job end to end: host arrays in, host arrays out, through Shard.predict (factorisation, H2D of the rank's block,
    # every predict launch, all_gather, D2H of all M results) -- args.c4_points per GPU
    Mj = args.c4_points * world
    Xj = c4_test_points(Mj)
    barrier()
    t0 = time.perf_counter()
    eng4.factorize(th4)
    muj, varj = shard.predict(eng4, Xj, epilogue=epi)
    barrier()
    tj = time.perf_counter() - t0
    rec4['e2e'] = {'value': Mj / tj, 'unit': 'pts/s', 'seconds': tj, 'points': Mj, 'points_per_gpu': args.c4_points,
                   'h2d_bytes_per_step': int(args.c4_points * 10 * 8), 'd2h_bytes_per_step': int(Mj * 16),
                   'frac': Mj * flops_predict(8192, 10) / tj / 1e12 / (p64 * world),
                   'api': 'engine.factorize + Shard.predict (NumPy [M,10] in, NumPy mean/var [M] out on every rank)',
                   'mean_range': [float(muj.min()), float(muj.max())], 'var_range': [float(varj.min()), float(varj.max())]}
    if cpu_legs:
        # CPU side of c4: oracle predict with L factorised once (kinder than the reference, which refactorises and
        # recompiles on every call), 2 blocks of 4096 points, vectorised GH epilogue; plus the reference's literal
        # per-point GH loop (gpmcmc.py:549-563) on 10^4 points
        from oracle import gp_oracle as go
        use_all_host_threads()
        spec4 = oracle_spec('c4')
        cores, api = blas_threads()
        t0 = time.perf_counter()
        _, _, L4 = go.predict_blocked(spec4, th4, X4, y4, X4[:8], block=8)
        t_fact = time.perf_counter() - t0
        Xc = Xj[:8192]
        rev = lambda v: (v - ab4[0]) / ab4[1]   # noqa: E731
        t0 = time.perf_counter()
        mu_c, var_c, _ = go.predict_blocked(spec4, th4, X4, y4, Xc, block=4096, L=L4)
        m_c, v_c = go.gh_stats(mu_c, var_c, rev, normvar=False)
        t_pred = time.perf_counter() - t0
        t0 = time.perf_counter()
        go.gh_stats_loop(np.resize(mu_c, 10000), np.resize(var_c, 10000), rev, normvar=False)
        t_loop = time.perf_counter() - t0
        rec4['cpu_baseline'] = {
            'value': 8192 / t_pred, 'unit': 'pts/s', 'cores': cores, 'kind': 'port', 'blas': api,
            'sample': '8192 of the test points in 2 blocks of 4096, L factorised once beforehand '
                      f'({t_fact:.1f} s, not counted), vectorised GH epilogue',
            'factorize_s': t_fact, 'reference_gh_loop_pts_per_s': 10000 / t_loop,
            # the job's first 8192 results against the oracle, in passing (the parity tests do this properly)
            'max_rel_dev_mean_vs_gpu': float(np.max(np.abs(muj[:8192] - m_c[:, 0]) / np.abs(m_c[:, 0]))),
            'max_rel_dev_var_vs_gpu': float(np.max(np.abs(varj[:8192] - v_c[:, 0]) / np.abs(v_c[:, 0])))}
    per['c4'] = rec4
    del muj, varj, Xj, g4, Xs

    # f3 / f4 on the c4 data (rank 0): rank-1 append vs refactorisation; inverse-problem potential
    if rank == 0:
        from andvaranaut_b200.xpost import InverseLikelihood, XPosterior
        import scipy.stats as st
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
        per['f3_append'] = {'metric': 'rank1_append_ms', 'value': t_app * 1e3, 'unit': 'ms', 'N': Na, 'bound': 'hbm',
                            'refactorize_ms': t_fact * 1e3, 'info': int(info_a[0]), 'algorithmic_bytes': 8 * Na * Na,
                            'achieved': 8 * Na * Na / t_app / 1e9, 'peak': hbm_peak, 'frac': 8 * Na * Na / t_app / 1e9 / hbm_peak,
                            'note': 'one new training point, hypers unchanged: avn_gp_append (4 launches, lower triangle of '
                                    'T read twice = 8 N^2 B) vs avn_gp_factorize of the enlarged set; host-timed incl. the info sync'}
        pot = InverseLikelihood(enga, lambda x: (x, np.ones_like(x)), np.array([0.3]), 1e-2, 0.0, 0.0)
        xpost = XPosterior([st.uniform(0, 1)] * 10, pot)
        zq = np.random.default_rng(7).normal(size=(256, 10))
        xpost.logp_dlogp(zq, True)
        t0 = time.perf_counter()
        for _ in range(3):
            xpost.logp_dlogp(zq, True)
        t_inv = (time.perf_counter() - t0) / 3
        per['f4_inverse'] = {'metric': 'inverse_logp_dlogp_evals_per_s', 'value': 256 / t_inv, 'unit': 'evals/s', 'N': enga.N,
                             'candidates_per_call': 256, 'ms_per_call': t_inv * 1e3,
                             'note': 'inverse_opt potential + gradient w.r.t. x for 256 candidate points (restarts / chains) '
                                     'per call: one avn_gp_predict_grad + host Schur term, host-timed'}
        del enga, pot, xpost
    del eng4
    torch.cuda.empty_cache()

    # ---- c5: BO acquisition at N=4096, d=12: EI over 4096 candidates (blocks of candidates sharded) ----------
    (kw5, X5, y5, th5), ab5, yopt5 = workload_c5()
    yrev5 = [(T.OP_AFFINE_CONST, -1, (ab5[0], ab5[1], 0.0, 0.0))]
    eng5 = GPEngine(**kw5, device=dev)
    eng5.set_data(X5, y5)
    info5 = eng5.factorize(th5)
    cand = c5_candidates()
    Mc = cand.shape[0]
    lo, hi, per5 = shard.bounds(Mc)
    cand_dev = torch.as_tensor(cand[lo:hi], device=dev)
    epi5 = GPEngine.make_epilogue(mode='EI', deg=8, EIopt='min', yopt=yopt5, yrev=yrev5)
    g5 = torch.empty(world * per5, 2, dtype=torch.float64, device=dev) if world > 1 else None
    ms5 = timed(lambda: shard.predict_dev(eng5, cand_dev, Mc, gathered=g5, epilogue=epi5), steps, warm=2)
    ms5_e2e = timed_host(lambda: shard.predict(eng5, cand, epilogue=epi5), steps)
    rec5 = {'metric': 'gp_predict_points_per_s', 'value': Mc / (ms5 * 1e-3), 'unit': 'pts/s', 'ms_per_step': ms5,
            'workload': 'c5: BO acquisition at N=4096 d=12 Matern52: expected improvement (GH(8), meanstd reversion) over '
                        '4096 LHC candidates per iteration',
            'scaling': 'strong', 'n_gpus': world, 'bound': 'tensor',
            'frac': Mc * flops_predict(4096, 12) / (ms5 * 1e-3) / 1e12 / (p64 * world), 'info': int(info5[0]),
            'api': 'Shard.predict_dev (row-split predict kernels: one wave of 296 CTAs)',
            'e2e': {'value': Mc / (ms5_e2e * 1e-3), 'unit': 'pts/s', 'ms_per_step': ms5_e2e,
                    'h2d_bytes_per_step': int((hi - lo) * 12 * 8), 'd2h_bytes_per_step': int(Mc * 16),
                    'api': 'Shard.predict (NumPy candidates in, NumPy EI / variance out)'}}
    if cpu_legs:
        from oracle import gp_oracle as go
        use_all_host_threads()
        spec5 = oracle_spec('c5')
        cores, api = blas_threads()
        _, _, L5 = go.predict_blocked(spec5, th5, X5, y5, X5[:8], block=8)
        rev5 = lambda v: (v - ab5[0]) / ab5[1]   # noqa: E731
        t0 = time.perf_counter()
        mu_c, var_c, _ = go.predict_blocked(spec5, th5, X5, y5, cand, block=4096, L=L5)
        go.gh_stats(mu_c, var_c, rev5, normvar=False, EI=True, EIopt='min', yopt=yopt5)
        t_pred = time.perf_counter() - t0
        rec5['cpu_baseline'] = {'value': Mc / t_pred, 'unit': 'pts/s', 'cores': cores, 'kind': 'port', 'blas': api,
                                'sample': 'all 4096 candidates, L factorised once beforehand (not counted), vectorised GH / EI'}
    per['c5'] = rec5
    del eng5
    torch.cuda.empty_cache()

    # ---- rank 0 only: the sequential drivers through the GPMCMC API (c1 tutorial, c2 as fit() runs it, c5 loop) ----
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import c1_tutorial_probe
        import c2_map_fit
        import c5_bo_probe
        per['c1'] = c1_tutorial_probe.run()
        per['c2_fit'] = c2_map_fit.run(cpu=cpu_legs)
        rec5['bo_loop'] = {'metric': 'bo_iterations_per_s',
                           'workload': '4096 LHC candidates -> EI -> argmax -> append -> warm-started MAP refit through '
                                       'GPMCMC, 3 iterations at each training-set size',
                           'sizes': [c5_bo_probe.run(n) for n in (256, 1024, 4096)]}


if __name__ == '__main__':
    main()

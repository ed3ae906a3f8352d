"""Golden oracle outputs ON THE BENCHMARK'S OWN INPUTS (bench.workload_c2/c3/c4/c5, bench.theta_cloud), so that the
timed workload is the verified workload.  Run in the build container (about two minutes):

    python tests/golden/make_bench_golden.py

  bench_c2.npz  ll + gradient of 8 rows of the 64-sample hyperparameter cloud bench.py times (N=2000, d=8, P=30)
  bench_c3.npz  ll + gradient of 8 of the 512 chains (N=1000, d=6)
  bench_c4.npz  N=8192, d=10: latent mean / variance and the GH(8)-reverted mean / variance (literal per-point loop of
                gpmcmc.py:545-569, meanstd reversion) at 512 random test points (the first 512 of the job's points)
  bench_c5.npz  N=4096, d=12: latent mean / variance and expected improvement (opt_type='min') + its variance for the
                4096 LHC candidates of one BO iteration
Every file carries a checksum of the regenerated training data, so a drifting generator is noticed.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402

C2_ROWS = [0, 9, 18, 27, 36, 45, 54, 63]
C3_ROWS = [0, 73, 146, 219, 292, 365, 438, 511]


def data_sum(X, y):
    return np.array([zlib.crc32(np.ascontiguousarray(X).view(np.uint8)), zlib.crc32(np.ascontiguousarray(y).view(np.uint8))],
                    dtype=np.int64)


def main():
    bench.use_all_host_threads()
    for name, wl, rows, B, seed in (('c2', bench.workload_c2(), C2_ROWS, 64, 202), ('c3', bench.workload_c3(), C3_ROWS, 512, 303)):
        kw, X, y, th = wl
        spec = bench.oracle_spec(name)
        thetas = bench.theta_cloud(th, B, seed=seed)[rows]
        res = [go.loglik(spec, t, X, y) for t in thetas]
        assert all(r.info == 0 for r in res)
        np.savez(os.path.join(HERE, f'bench_{name}.npz'), rows=np.array(rows), thetas=thetas, data_sum=data_sum(X, y),
                 ll=np.array([r.ll for r in res]), grad=np.array([r.grad for r in res]))
        print(name, 'll', [round(r.ll, 3) for r in res], flush=True)
    # c4
    (kw, X, y, th), ab = bench.workload_c4()
    Xs = bench.c4_test_points(512)
    mu, var, _ = go.predict_blocked(bench.oracle_spec('c4'), th, X, y, Xs)
    rev = lambda v: (v - ab[0]) / ab[1]   # noqa: E731   affine.rev, transform.py:208-221
    m, v = go.gh_stats_loop(mu, var, rev, normvar=False, deg=8)
    np.savez(os.path.join(HERE, 'bench_c4.npz'), Xs=Xs, theta=th, ab=np.array(ab), data_sum=data_sum(X, y), mu=mu, var=var,
             gh_mean=m[:, 0], gh_var=v[:, 0])
    print('c4 mu', mu[:3], 'var', var[:3], 'gh', m[:3, 0], v[:3, 0], flush=True)
    # c5
    (kw, X, y, th), ab, yopt = bench.workload_c5()
    cand = bench.c5_candidates()
    mu, var, _ = go.predict_blocked(bench.oracle_spec('c5'), th, X, y, cand)
    rev = lambda v: (v - ab[0]) / ab[1]   # noqa: E731
    ei, eiv = go.gh_stats_loop(mu, var, rev, normvar=False, deg=8, EI=True, EIopt='min', yopt=yopt)
    m, v = go.gh_stats_loop(mu, var, rev, normvar=True, deg=8)
    # the running optimum makes EI zero at most candidates (2 of 4096 non-zero here): a second threshold inside the
    # predicted range (opt_type='max' against the median prediction) exercises the positive branch at about half of them
    yopt2 = float(np.median(m))
    ei2, ei2v = go.gh_stats_loop(mu, var, rev, normvar=False, deg=8, EI=True, EIopt='max', yopt=yopt2)
    np.savez(os.path.join(HERE, 'bench_c5.npz'), theta=th, ab=np.array(ab), yopt=np.array(yopt), data_sum=data_sum(X, y),
             cand_sum=data_sum(cand, cand[:, 0]), mu=mu, var=var, ei=ei[:, 0], ei_var=eiv[:, 0], gh_mean=m[:, 0],
             gh_normvar=v[:, 0], yopt2=np.array(yopt2), ei2=ei2[:, 0], ei2_var=ei2v[:, 0])
    print('ei2 nonzero', int((ei2 > 0).sum()))
    print('c5 EI max', ei.max(), 'nonzero', int((ei > 0).sum()), 'var range', var.min(), var.max(), flush=True)


if __name__ == '__main__':
    main()

"""CPU oracle for the andvaranaut GP inner loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``andvaranaut_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs use it, and there only as the checker or
as the timed CPU baseline -- never as the product path.

PARITY UNPINNED: the reference (andrew-angus/andvaranaut) delegates all GP
arithmetic to PyMC <= 5.9.2 / PyTensor / SciPy LAPACK, none of which can be
installed in the build container, and it ships no tests and no reproducible
golden vectors for the GP path (its LHC sampler drops the seed,
``andvaranaut/lhc.py:40-43``).  The formulas here restate PyMC 5.9.x
(``pymc/gp/cov.py``, ``pymc/gp/gp.py``, ``pymc/gp/util.py``,
``pymc/distributions/multivariate.py``, ``pymc/tuning/starting.py``) at the
reference's call sites (``andvaranaut/gpmcmc.py:185-401, 522-598``).  Only the
transform half is pinned: ``oracle/warp_oracle.py`` is checked against the
reference's own ``andvaranaut/transform.py`` (imported with a stubbed
``pytensor``) through the fixtures in ``tests/golden/`` and against the
tutorial's recorded known answers (``tutorial/tutorial.ipynb:362-369``).
"""

#!/usr/bin/env python
"""Generates tests/golden/ref_eigcg.npz from the reference's own incremental eigCG (generic_ks/inc_eigcg.c compiled
unmodified into oracle/_ref/libmilcref.so by oracle/build_ref.sh, driven through oracle/ref_harness/eigcg_harness.c):
a sequence of ks_inc_eigCG_parity solves on a seeded 4x4x4x8 lattice, then calc_eigenpairs.
    python tests/golden/make_golden_eigcg.py
Inputs are regenerated from the seeds by the tests (milc_qcd_b200.fields), only results are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from milc_qcd_b200 import fields as F  # noqa: E402

DIMS, MASS, RESID, M, NVECS, NMAX, NSOLVES = (4, 4, 4, 8), 0.05, 1e-10, 40, 6, 24, 5
EVEN = 2

fat, lng = F.make_links(DIMS, seed=1234)
ref = pyoracle.MilcRef(DIMS, "")
assert ref.has_eigcg, "oracle/_ref was built without inc_eigcg.c (no LAPACK found)"
ref.set_links(fat, lng)
ref.inc_eigcg_init(M, NVECS, NMAX)
iters, ncurr, sols = [], [], []
for s in range(NSOLVES):
    b = F.make_source(DIMS, seed=2000 + s, parity=EVEN)
    x = np.zeros_like(b)
    it, q, n = ref.inc_eigcg(b, x, MASS, EVEN, 2000, 5, RESID)
    assert q["converged"] == 1
    iters.append(it)
    ncurr.append(n)
    sols.append(x[: x.shape[0] // 2].copy())
val, vec = ref.eigcg_pairs(EVEN, ncurr[-1])
# the single-solve form with a fixed number of iterations
b = F.make_source(DIMS, seed=1235, parity=EVEN)
x = np.zeros_like(b)
it1, val1, vec1, q1 = ref.eigcg(b, x, MASS, EVEN, 2000, 5, RESID, M, NVECS)
np.savez_compressed(os.path.join(HERE, "ref_eigcg.npz"), dims=np.array(DIMS), mass=MASS, resid=RESID, m=M, nvecs=NVECS, nmax=NMAX,
                    iters=np.array(iters), ncurr=np.array(ncurr), sols=np.array(sols), eigval=val,
                    eigvec_even=vec[:, : vec.shape[1] // 2].astype(np.float32),
                    single_iters=it1, single_eigval=val1, single_sol=x[: x.shape[0] // 2])
print("iterations per solve", iters, "accumulated", ncurr)
print("lowest Ritz values", val[:8])

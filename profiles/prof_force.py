"""Profiling target for the fermion force: one b200ks_hisq_force call (3 terms) on the bench lattice."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

dims = (32, 32, 32, 64)
V = int(np.prod(dims))
ctx = api.Context(dims)
ctx.hisq_links_time(1234, 1)
U, Vl, W = ctx.hisq_links_fetch(0), ctx.hisq_links_fetch(1), ctx.hisq_links_fetch(2)
rng = np.random.default_rng(1)
X = [rng.standard_normal((V, 3, 2)) for _ in range(3)]
mom = ctx.hisq_force(U, Vl, W, X, [0.3, 0.5, 0.7], 0.02)
print("done", float(np.abs(mom).max()))
ctx.close()

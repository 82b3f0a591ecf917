"""Profiling target for the low-mode deflation (not yet run): three b200ks_deflate_dev calls with 64 resident
synthetic vectors on the bench lattice.  Launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_deflate.csv \
        python profiles/prof_deflate.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

dims = (32, 32, 32, 64)
nvecs = int(os.environ.get("NVECS", "64"))
ctx = api.Context(dims)
hs = []
for j in range(nvecs):
    h = ctx.vec_create()
    ctx.vec_gaussian(h, 3, 1000 + j)
    hs.append(h)
ctx.eig_set(hs, list(np.linspace(1e-4, 1e-2, nvecs)), use_in_uml=False)
vs, vd = ctx.vec_create(), ctx.vec_create()
ctx.vec_gaussian(vs, 3, 7)
ctx.vec_zero(vd, 3)
for _ in range(3):
    ctx.deflate_dev(vs, vd, 0.05, 2)
print("done", ctx.vec_norm2(vd, 2))
ctx.close()

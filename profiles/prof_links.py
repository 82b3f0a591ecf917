"""Profiling target for the fermion-link construction: one HISQ chain on the bench lattice
(device-generated thin links)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

ctx = api.Context((32, 32, 32, 64))
ms, nsvd = ctx.hisq_links_time(1234, 1)
print("chain ms", ms, "svd links", nsvd)
ctx.close()

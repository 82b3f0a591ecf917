"""Profiling target: the 16-bit stencil at the 8-GPU local volume of BASELINE configs[3] (64x64x32x24), unpartitioned or
(B200KS_FORCE_PARTITION=zt) with z and t partitioned and the GPU as its own neighbour -- the same sites either way."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

ctx = api.Context((64, 64, 32, 24), grid=(1, 1, 1, 1), rank=0, nranks=1)
ctx.links_synthetic(1234)
print("halo mode", ctx.halo_mode(), "ms per launch", ctx.dslash_time(0, 2, 12))
ctx.close()

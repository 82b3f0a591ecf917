"""Profiling target: mixed CG on the 8-GPU local volume of BASELINE configs[3] (64x64x32x24) with z and t
partitioned and the GPU as its own neighbour (B200KS_FORCE_PARTITION=zt): the launch list of one rank of
an 8-GPU run -- push kernel, single-launch interior/boundary stencil, finish kernels -- without the
other seven (a real multi-rank run cannot go under ncu: kernels that wait for a peer's kernel would be
serialised behind it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

ctx = api.Context((64, 64, 32, 24), grid=(1, 1, 1, 1), rank=0, nranks=1)
ctx.links_synthetic(1234)
vb, vx = ctx.vec_create(), ctx.vec_create()
ctx.vec_gaussian(vb, 2, 5678)
it, res = ctx.congrad_dev(vb, vx, 0.05, 2, 60, 1, 1e-10, mixed_precision=2)
print("iters", it, res["final_rsq"])
ctx.close()

"""Profiling target that launches every stencil variant of the single-GPU solvers exactly once or
twice (a --set full capture costs ~8 s per launch): plain stencils through b200ks_dslash_dev /
b200ks_dslash_block_dev, the fused-dots variants through solves capped at one iteration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

EVEN, ODD = 2, 1
dims = (32, 32, 32, 64)
ctx = api.Context(dims)
ctx.links_synthetic(1234, 0)
vb = [ctx.vec_create() for _ in range(4)]
vx = [ctx.vec_create() for _ in range(4)]
for k in range(4):
    ctx.vec_gaussian(vb[k], EVEN, 5678 + 101 * k)
for prec in (0, 1, 2):
    ctx.dslash_dev(vb[0], vx[0], ODD, prec)
for prec in (1, 2):
    ctx.dslash_block_dev(vb, vx, ODD, prec)
ctx.vec_zero(vx[0], EVEN)
print("single mixed 2", ctx.congrad_dev(vb[0], vx[0], 0.05, EVEN, 1, 1, 1e-10, mixed_precision=2)[0])
for mixed in (1, 0):
    for k in range(4):
        ctx.vec_zero(vx[k], EVEN)
    print("block", mixed, ctx.congrad_block_dev(vb, vx, 0.05, EVEN, 1, 1, 1e-10, mixed_precision=mixed)[0])
ctx.close()

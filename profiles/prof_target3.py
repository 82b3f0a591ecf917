"""Second profiling target: the fused-dots stencil variants, one launch each (solves capped at
max_iter = 2: one true-residual evaluation, one inner iteration, one final evaluation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

EVEN, ODD = 2, 1
ctx = api.Context((32, 32, 32, 64))
ctx.links_synthetic(1234, 0)
vb = [ctx.vec_create() for _ in range(4)]
vx = [ctx.vec_create() for _ in range(4)]
for k in range(4):
    ctx.vec_gaussian(vb[k], EVEN, 5678 + 101 * k)
ctx.dslash_dev(vb[0], vx[0], ODD, 1)
for mixed in (2, 1):
    ctx.vec_zero(vx[0], EVEN)
    print("single", mixed, ctx.congrad_dev(vb[0], vx[0], 0.05, EVEN, 2, 1, 1e-10, mixed_precision=mixed, check_interval=1)[0])
for mixed in (1, 0):
    for k in range(4):
        ctx.vec_zero(vx[k], EVEN)
    print("block", mixed, ctx.congrad_block_dev(vb, vx, 0.05, EVEN, 2, 1, 1e-10, mixed_precision=mixed, check_interval=1)[0])
ctx.close()

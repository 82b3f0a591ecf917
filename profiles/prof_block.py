"""Profiling target for the block (multi-right-hand-side) CG: 4 sources on the bench lattice,
capped at 24 iterations, mixed precision (float K-wide stencil) then pure double."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

EVEN = 2
dims = tuple(int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 32, 32, 64)
nsrc = int(os.environ.get("NSRC", "4"))
ctx = api.Context(dims)
ctx.links_synthetic(1234, 0)
vb = [ctx.vec_create() for _ in range(nsrc)]
vx = [ctx.vec_create() for _ in range(nsrc)]
for k in range(nsrc):
    ctx.vec_gaussian(vb[k], EVEN, 5678 + 101 * k)
for mixed in (1, 0):
    for k in range(nsrc):
        ctx.vec_zero(vx[k], EVEN)
    it, res = ctx.congrad_block_dev(vb, vx, 0.05, EVEN, 24, 1, 1e-10, mixed_precision=mixed)
    print("mixed", mixed, "iters", it, "rsq", [r["final_rsq"] for r in res])
ctx.close()

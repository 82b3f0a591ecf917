"""Short profiling target: one capped mixed-precision CG and one capped double CG on the bench
workload (32^3x64 synthetic), so that `ncu --set full -k regex:dslash_kernel -c 6` sees the
double and the single-precision stencil variants within the first few launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

EVEN = 2
dims = tuple(int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 32, 32, 64)
ctx = api.Context(dims)
ctx.links_synthetic(1234, int(os.environ.get("LONG_RECON", "0")))
print("long links: %d complex per link, misfit %.2e" % ctx.long_link_info())
vb, vx = ctx.vec_create(), ctx.vec_create()
ctx.vec_gaussian(vb, EVEN, 5678)
for mixed in (2, 1, 0):
    ctx.vec_zero(vx, EVEN)
    it, res = ctx.congrad_dev(vb, vx, 0.05, EVEN, 24, 1, 1e-10, mixed_precision=mixed)  # 24 iterations: several launches of every stencil variant
    print("mixed", mixed, "iters", it, "rsq", res["final_rsq"])
ctx.close()

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api
ctx = api.Context((32, 32, 32, 64))
ctx.links_synthetic(1234)
vb, vx = ctx.vec_create(), ctx.vec_create()
ctx.vec_gaussian(vb, 2, 5678)
print("dslash ms f16/f32:", ctx.dslash_time(0, 2, 100), ctx.dslash_time(1, 2, 100))
for mixed in (2, 1):
    for rep in range(2):
        ctx.vec_zero(vx, 2)
        it, res = ctx.congrad_dev(vb, vx, 0.05, 2, 2000, 10, 1e-10, mixed_precision=mixed)
    print("mixed", mixed, "iters", it, "us/iter", 1e6 * res["device_seconds"] / it, "restarts", res["final_restart"], "rsq", res["final_rsq"])
ctx.close()

"""One-GPU probe of the multi-GPU stencil path: the 8-GPU local volume of BASELINE configs[3]
(64x64x32x24) with z and t treated as partitioned (the GPU is its own neighbour), against the
same volume unpartitioned.  The difference is the cost of the interior/exterior split, the push
kernel and the flag wait -- everything except NVLink time of flight."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

EVEN = 2
dims = tuple(int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 64, 32, 24)
out = {}
for force in ("", "zt"):
    if force:
        os.environ["B200KS_FORCE_PARTITION"] = force
    else:
        os.environ.pop("B200KS_FORCE_PARTITION", None)
    ctx = api.Context(dims, grid=(1, 1, 1, 1), rank=0, nranks=1)
    ctx.links_synthetic(1234)
    vb, vx = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vb, EVEN, 5678)
    row = {"halo_mode": ctx.halo_mode()}
    for prec in (2, 1, 0):
        row["dslash_ms_f%d" % (32 * prec if prec else 16)] = ctx.dslash_time(prec, EVEN, 200)
    for mixed in (0, 1, 2):
        ctx.vec_zero(vx, EVEN)
        it, res = ctx.congrad_dev(vb, vx, 0.05, EVEN, 2000, 10, 1e-10, mixed_precision=mixed)
        ctx.vec_zero(vx, EVEN)
        it, res = ctx.congrad_dev(vb, vx, 0.05, EVEN, 2000, 10, 1e-10, mixed_precision=mixed)
        row["cg_mixed%d" % mixed] = {"iters": it, "seconds": res["device_seconds"], "us_per_iter": 1e6 * res["device_seconds"] / it}
    row["fused_push"] = os.environ.get("B200KS_FUSED_PUSH", "1")
    row["pdl"] = os.environ.get("B200KS_PDL", "1")
    out[force or "none"] = row
    ctx.close()
print(json.dumps(out, indent=1))

"""Breakdown of the partitioned stencil on one GPU (self-neighbour): B200KS_DEBUG_SKIP masks out
the push+wait (1), the interior launch (2), the boundary launch (4)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    sys.path.insert(0, ROOT)
    from milc_qcd_b200 import api
    dims = (64, 64, 32, 24)
    ctx = api.Context(dims, grid=(1, 1, 1, 1), rank=0, nranks=1)
    ctx.links_synthetic(1234)
    print(json.dumps({"f64": ctx.dslash_time(2, 2, 200), "f32": ctx.dslash_time(1, 2, 200)}))
    ctx.close()
    sys.exit(0)
out = {}
for force in ("t", "zt"):
    for skip in (0, 1, 2, 4, 3, 5, 6):
        env = dict(os.environ, B200KS_FORCE_PARTITION=force, B200KS_DEBUG_SKIP=str(skip))
        r = subprocess.run([sys.executable, __file__, "worker"], env=env, capture_output=True, text=True)
        out["%s skip=%d" % (force, skip)] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]
print(json.dumps(out, indent=1))

"""Times the 16-bit stencil and the mixed_precision-2 CG for several builds of libb200ks
(profiles/variants/*.so, built with different -D flags; see profiles/variants/README).  Each
build runs in its own process (B200KS_LIB selects the library).  Output: one JSON per build.

    python profiles/variant_probe.py            # driver: all variants
    python profiles/variant_probe.py --one      # worker: the library named by B200KS_LIB
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    from milc_qcd_b200 import api
    dims = (32, 32, 32, 64)
    ctx = api.Context(dims)
    ctx.links_synthetic(1234, 0)
    out = {"lib": os.environ.get("B200KS_LIB", "default")}
    for prec in (0, 1):
        ts = [ctx.dslash_time(prec, 2, 100) for _ in range(3)]
        out["dslash_ms_prec%d" % prec] = min(ts)
    if "probe_" in out["lib"]:      # diagnostic builds (no math / no link loads): stencil timings only
        for prec in (0,):
            out["dslash_ms_long_prec%d" % prec] = ctx.dslash_time(prec, 2, 3000)
        print("VARIANT " + json.dumps(out))
        ctx.close()
        return
    out["dslash_ms_long_prec0"] = ctx.dslash_time(0, 2, 3000)
    for prec in (1, 2):
        for k in (3, 4):
            out["dslash_block_ms_prec%d_k%d" % (prec, k)] = min(ctx.dslash_block_time(prec, k, 2, 100) for _ in range(2))
    vb, vx = ctx.vec_create(), ctx.vec_create()
    ctx.vec_gaussian(vb, 2, 5678)
    best = None
    for rep in range(3):
        ctx.vec_zero(vx, 2)
        it, res = ctx.congrad_dev(vb, vx, 0.05, 2, 2000, 10, 1e-10, mixed_precision=2)
        if best is None or res["device_seconds"] < best[1]:
            best = (it, res["device_seconds"], res["final_rsq"])
    out["cg_mixed2"] = {"iters": best[0], "seconds": best[1], "final_rsq": best[2]}
    for mixed in (0, 1):
        best = None
        for rep in range(2):
            ctx.vec_zero(vx, 2)
            it, res = ctx.congrad_dev(vb, vx, 0.05, 2, 2000, 10, 1e-10, mixed_precision=mixed)
            if best is None or res["device_seconds"] < best[1]:
                best = (it, res["device_seconds"])
        out["cg_mixed%d" % mixed] = {"iters": best[0], "seconds": best[1]}
    vbs = [ctx.vec_create() for _ in range(4)]
    vxs = [ctx.vec_create() for _ in range(4)]
    for k in range(4):
        ctx.vec_gaussian(vbs[k], 2, 5678 + 101 * k)
    for mixed in (0, 1):
        best = None
        for rep in range(2):
            for k in range(4):
                ctx.vec_zero(vxs[k], 2)
            it, res = ctx.congrad_block_dev(vbs, vxs, 0.05, 2, 2000, 10, 1e-10, mixed_precision=mixed)
            if best is None or res[0]["device_seconds"] < best[1]:
                best = (it, res[0]["device_seconds"])
        out["block4_mixed%d" % mixed] = {"iters": best[0], "seconds": best[1]}
    print("VARIANT " + json.dumps(out))
    ctx.close()


def main():
    libs = [None] + sorted(glob.glob(os.path.join(ROOT, "profiles", "variants", "*.so")))
    rows = []
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["B200KS_LIB"] = lib
        else:
            env.pop("B200KS_LIB", None)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True,
                           timeout=300)
        for line in p.stdout.splitlines():
            if line.startswith("VARIANT "):
                rows.append(json.loads(line[8:]))
        if p.returncode != 0:
            rows.append({"lib": lib, "error": p.stderr[-400:]})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "variant_probe.json"), "w") as f:
        json.dump(rows, f, indent=1)
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    one() if "--one" in sys.argv else main()

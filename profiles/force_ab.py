"""A/B of the launch variants of the link construction and the fermion force on the bench lattice (one process, one GPU):
B200KS_SITE_ORDER (parities interleaved CTA by CTA), B200KS_FORCE_SPLIT (backward staple passes fused / four kernels / two
roles of one kernel at 128 or 168 registers), B200KS_FORCE_OVERLAP (V and U uploaded while the W-level chain runs).
The library reads the switches on every call.  Prints one JSON object: link-chain milliseconds per order, seconds per
force call per variant (9 terms, host buffers in and out, like bench.py --workload force) and the largest deviation of
every variant's momenta from the baseline variant's.

    python profiles/force_ab.py                 # all variants
    python profiles/force_ab.py --only 0,0,0:1,0,0    # one 3-term call per listed variant (profiling target for ncu)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from milc_qcd_b200 import api  # noqa: E402

KEYS = ("B200KS_SITE_ORDER", "B200KS_FORCE_SPLIT", "B200KS_FORCE_OVERLAP")
VARIANTS = [(0, 0, 0), (1, 0, 0), (0, 2, 0), (1, 2, 0), (1, 3, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (1, 2, 1), (1, 3, 1)]


def set_variant(v):
    for k, x in zip(KEYS, v):
        os.environ[k] = str(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--lattice", type=int, nargs=4, default=[32, 32, 32, 64])
    ap.add_argument("--calls", type=int, default=3)
    args = ap.parse_args()
    dims = tuple(args.lattice)
    V = int(np.prod(dims))
    ctx = api.Context(dims)
    out = {"lattice": list(dims), "switches": list(KEYS)}
    if args.only:   # "o,f,v" or several of them separated by ":" -- one 3-term force call each, in this order
        vs = [tuple(int(x) for x in w.split(",")) for w in args.only.split(":")]
        set_variant(vs[0])
        ms, _ = ctx.hisq_links_time(1234, 1)
        U, Vl, W = ctx.hisq_links_fetch(0), ctx.hisq_links_fetch(1), ctx.hisq_links_fetch(2)
        rng = np.random.default_rng(1)
        X = [rng.standard_normal((V, 3, 2)) for _ in range(3)]
        for v in vs:
            set_variant(v)
            mom = ctx.hisq_force(U, Vl, W, X, [0.3, 0.5, 0.7], 0.02)
            print(json.dumps({"variant": v, "links_chain_ms": ms, "mom_max": float(np.abs(mom).max())}))
        ctx.close()
        return
    links = {}
    for order in (0, 1, 0, 1):
        set_variant((order, 0, 0))
        ms, nsvd = ctx.hisq_links_time(1234, 3)
        links.setdefault(str(order), []).append(ms)
    out["links_chain_ms_by_site_order"] = links
    U, Vl, W = ctx.hisq_links_fetch(0), ctx.hisq_links_fetch(1), ctx.hisq_links_fetch(2)
    rng = np.random.default_rng(77)
    nterms = 9
    X = [rng.standard_normal((V, 3, 2)) for _ in range(nterms)]
    res = np.linspace(0.2, 1.0, nterms)
    base = None
    rows = []
    for v in VARIANTS:
        set_variant(v)
        times = []
        for _ in range(args.calls):
            t0 = time.perf_counter()
            mom = ctx.hisq_force(U, Vl, W, X, res, 0.02)
            times.append(time.perf_counter() - t0)
        if base is None:
            base = mom
        rows.append({"site_order": v[0], "form": v[1], "overlap": v[2], "seconds_per_call": min(times), "all_calls_s": times,
                     "max_abs_dev_from_baseline": float(np.abs(mom - base).max()), "mom_max": float(np.abs(mom).max())})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    out["force"] = rows
    out["nterms"] = nterms
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_hisq_links.npz from the REFERENCE ITSELF.

    python tests/golden/make_golden_links.py      # build container: needs oracle/_ref/libmilcref.so

Input: seeded thin SU(3) links with KS phases and the antiperiodic boundary folded in
(milc_qcd_b200.fields.make_thin_links) on a 4x4x4x6 lattice, one smooth field (spread 0.4) and
one close to strong coupling (spread 5: a few links take the reference's SVD branch).  Outputs:
the reference's own create_hisq_links_milc chain -- V (fat7), W (U(3) projection), fat, long --
from its compiled sources (generic_ks/fermion_links_hisq_load_milc.c, fermion_links_fn_load_milc.c,
generic/general_staple.c, generic_ks/su3_mat_op.c, path tables of hisq_u3_action.h).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from milc_qcd_b200 import fields as F  # noqa: E402
from oracle.pyoracle import MilcRef  # noqa: E402


def main():
    dims = (4, 4, 4, 6)
    ref = MilcRef(dims)
    out = {"dims": np.array(dims)}
    for tag, spread in (("smooth", 0.4), ("rough", 5.0)):
        U = F.make_thin_links(dims, seed=4321, spread=spread)
        r = ref.hisq_links(U)
        out[tag + "_spread"] = spread
        out[tag + "_nsvd"] = r["nsvd"]
        out["coeffs"] = r["coeffs"]
        for k in ("V", "W", "fat", "lng"):
            out[tag + "_" + k] = r[k]
        print(tag, "SVD branch:", r["nsvd"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hisq_links.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

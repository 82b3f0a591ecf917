#!/usr/bin/env python
"""Generates tests/golden/ref_meson.npz from the reference's own ks_meson_cont_mom (generic_ks/ks_meson_mom.c,
spin_taste_ops.c and generic_wilson/gammas.c compiled unmodified into oracle/_ref/libmilcref.so by oracle/build_ref.sh,
driven through oracle/ref_harness/meson_harness.c) on a seeded 4x6x4x8 lattice:
  prop_local    all LOCAL sink operators of tests/meson_case.py (site signs), 5 momenta with mixed reflection parities
  prop_shifted  link-shift operators (one-link with APE links, FN vector currents, a gamma-gamma one-link operator)
  op_fields     what spin_taste_op_fn makes of the two propagators for those operators (the fields a MILC build hands to
                the device contraction with spin = -1), so that the check also runs where /root/reference is absent
    python tests/golden/make_golden_meson.py
Inputs are regenerated from the seeds by the tests, only results are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle  # noqa: E402
from milc_qcd_b200 import fields as F, meson as M  # noqa: E402
import meson_case as K  # noqa: E402

ref = pyoracle.MilcRef(K.DIMS, "")
assert ref.has_meson
fat, lng = F.make_links(K.DIMS, seed=11)
ref.set_links(fat, lng)
ref.set_ape_links(F.make_links(K.DIMS, seed=77)[0])
s1, s2 = K.sources()
index_of = {nm: ref.spin_taste_index(nm) for nm in K.LOCAL + K.SHIFTED}
assert all(v >= 0 for v in index_of.values()), index_of
out = {}
for tag, names in (("local", K.LOCAL), ("shifted", K.SHIFTED)):
    st, pi, ph, fa, ci, ct = K.table(index_of, names)
    out["prop_" + tag] = ref.meson_cont_mom(s1, s2, K.MOM, K.PAR, st, pi, ph, fa, ci, K.NPROP, K.R0)
ops = {}
for nm in K.SHIFTED:
    i = index_of[nm]
    if M.is_rhosfn(i):
        ops[nm + "/b1"] = ref.spin_taste_op(M.backward_index(i), K.R0, s1)
        ops[nm + "/f2"] = ref.spin_taste_op(M.forward_index(i), K.R0, s2)
    else:
        ops[nm + "/1"] = ref.spin_taste_op(i, K.R0, s1)
np.savez_compressed(os.path.join(HERE, "ref_meson.npz"), names=np.array(list(index_of)), index=np.array(list(index_of.values())),
                    op_keys=np.array(list(ops)), op_fields=np.array(list(ops.values())), **out)
print({k: float(np.abs(v).max()) for k, v in out.items()}, len(ops), "operator fields")

"""Generates tests/golden/ref_hisq_force.npz from the REFERENCE ITSELF.

    python tests/golden/make_golden_force.py      # build container: needs oracle/_ref/libmilcref.so

Input: seeded thin links (phases in) on a 4^4 lattice, two "solution" vectors X_j on the even
sites with their odd sites filled by D X_j through the HISQ links built from those thin links
(what ks_imp_rhmc/update_h_rhmc.c:75-86 hands to the force), residues (0.7, -0.3), eps = 1.
Output: the momentum update of the reference's eo_fermion_force_multi
(generic_ks/fermion_force_hisq_multi.c:170-216, wrapper_mx path, ks_imp_rhmc's build flags) as
anti_hermitmat arrays.  A second file, ref_hisq_force_rough.npz, holds the same for links rough enough
to trip the reference's eigenvalue filter and SVD branches; a third, ref_hisq_force_naik.npz, five terms in
three Naik-epsilon classes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from milc_qcd_b200 import fields as F  # noqa: E402
from oracle.pyoracle import MilcRef, LinksOracle, Oracle, ODD  # noqa: E402


def main():
    dims = (4, 4, 4, 4)
    V = int(np.prod(dims))
    h = V // 2
    ref, lo, o = MilcRef(dims), LinksOracle(), Oracle()
    U = F.make_thin_links(dims, seed=11, spread=0.5)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((2, V, 3, 2))
    X[:, h:] = 0
    residues = np.array([0.7, -0.3])
    links = lo.hisq_links(dims, U)
    for j in range(2):
        X[j, h:] = o.dslash(dims, links["fat"], links["lng"], X[j], ODD)[h:]
    mom, n = ref.hisq_force(U, X, residues, 1.0)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hisq_force.npz")
    np.savez_compressed(path, dims=np.array(dims), U=U, multi_x=X, residues=residues, eps=1.0, mom=mom, nsvd=n)
    print("wrote", path, os.path.getsize(path), "bytes; |mom|max", np.abs(mom).max())

    # rough links: some eigenvalues of Q = V^+ V fall below HISQ_FORCE_FILTER (5e-5) and below the
    # SVD thresholds, so the reference takes its filter and SVD branches (count = nsvd)
    U = F.make_thin_links(dims, seed=11, spread=0.8)
    links = lo.hisq_links(dims, U)
    X[:, h:] = 0
    for j in range(2):
        X[j, h:] = o.dslash(dims, links["fat"], links["lng"], X[j], ODD)[h:]
    mom, n = ref.hisq_force(U, X, residues, 1.0)
    assert n > 0
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hisq_force_rough.npz")
    np.savez_compressed(path, dims=np.array(dims), U=U, multi_x=X, residues=residues, eps=1.0, mom=mom, nsvd=n)
    print("wrote", path, os.path.getsize(path), "bytes; |mom|max", np.abs(mom).max(), "filter+svd links", n)

    # several Naik epsilons (the charm quark's mass-dependent correction): three classes of 2, 1 and 2 terms;
    # class k was solved with fat_k = fat_0 + eps_k/8 W, lng_k = (1 + eps_k) lng_0
    U = F.make_thin_links(dims, seed=11, spread=0.5)
    links = lo.hisq_links(dims, U)
    n_orders, eps_naik, cls = [2, 1, 2], [0.0, -0.0358, -0.2297], [0, 0, 1, 2, 2]
    X = rng.standard_normal((5, V, 3, 2))
    X[:, h:] = 0
    residues = np.array([0.7, -0.3, 0.45, 0.2, -0.6])
    for j in range(5):
        e = eps_naik[cls[j]]
        X[j, h:] = o.dslash(dims, links["fat"] + e * lo.NAIK_TABLE[0] * links["W"], (1 + e) * links["lng"], X[j], ODD)[h:]
    mom, n, fl = ref.hisq_force_naik(U, X, residues, n_orders, eps_naik, 1.0, want_links=True)
    for k, e in enumerate(eps_naik):   # the reference's links of every class are what the comment above says
        assert np.abs(fl[k, 0] - (links["fat"] + e * lo.NAIK_TABLE[0] * links["W"])).max() < 1e-12
        assert np.abs(fl[k, 1] - (1 + e) * links["lng"]).max() < 1e-12
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hisq_force_naik.npz")
    np.savez_compressed(path, dims=np.array(dims), U=U, multi_x=X, residues=residues, eps=1.0, mom=mom, nsvd=n,
                        n_orders=np.array(n_orders), eps_naik=np.array(eps_naik))
    print("wrote", path, os.path.getsize(path), "bytes; |mom|max", np.abs(mom).max())


if __name__ == "__main__":
    main()

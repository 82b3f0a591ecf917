"""Generates tests/golden/ref_uml_deflated.npz from the REFERENCE ITSELF.

    python tests/golden/make_golden_deflate.py      # build container: needs oracle/_ref/libmilcref.so

The reference's mat_invert_uml_field with qic->deflate = 1 (generic_ks/mat_invert.c:131-183,328-402) on a 4^4
lattice with the k lowest exact eigenpairs of -D_eo D_oe (dense diagonalisation of the oracle's stencil; odd
sites filled with D v / sqrt(lambda)), k = 8, 48 and all 384: iteration counts and solutions.  With every mode
in the set the trial solution is exact and both CGs stop after their first true-residual check.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from milc_qcd_b200 import fields as F  # noqa: E402
from oracle.pyoracle import MilcRef, Oracle, EVEN, ODD, EVENANDODD  # noqa: E402

DIMS, LINK_SEED, SRC_SEED, MASS, NITER, NRESTART, RESID = (4, 4, 4, 4), 77, 200, 0.02, 500, 5, 1e-10


def low_modes(o, dims, fat, lng):
    """All eigenpairs of -D_eo D_oe on the even sites, as (eigval, eigvec[(n, V, 3, 2)]) with the odd sites
    filled with D v / sqrt(lambda) (an orthonormal eigenbasis of -D_oe D_eo)."""
    V = int(np.prod(dims))
    h = V // 2
    n = 3 * h
    A = np.zeros((n, n), complex)
    for k in range(n):
        e = np.zeros((V, 3, 2))
        e[k // 3, k % 3, 0] = 1
        d1 = o.dslash(dims, fat, lng, e, ODD)
        d1[:h] = 0
        d2 = o.dslash(dims, fat, lng, d1, EVEN)
        A[:, k] = -(d2[:h, :, 0] + 1j * d2[:h, :, 1]).reshape(n)
    lam, vec = np.linalg.eigh(0.5 * (A + A.conj().T))
    ev = np.zeros((n, V, 3, 2))
    for j in range(n):
        ve = vec[:, j].reshape(h, 3)
        ev[j, :h, :, 0], ev[j, :h, :, 1] = ve.real, ve.imag
        ev[j, h:] = o.dslash(dims, fat, lng, ev[j], ODD)[h:] / np.sqrt(lam[j])
    return lam, ev


def main():
    fat, lng = F.make_links(DIMS, seed=LINK_SEED)
    o, r = Oracle(), MilcRef(DIMS)
    r.set_links(fat, lng)
    lam, ev = low_modes(o, DIMS, fat, lng)
    src = F.make_source(DIMS, seed=SRC_SEED, parity=EVENANDODD)
    ks, iters, sols = [8, 48, len(lam)], [], []
    for k in ks:
        dst = np.zeros_like(src)
        it, q = r.mat_invert_uml_deflated(src, dst, MASS, NITER, NRESTART, RESID, ev[:k], lam[:k])
        assert q["converged"] == 1
        iters.append(it)
        sols.append(dst)
    dst = np.zeros_like(src)
    it0, _ = r.mat_invert_uml(src[None], dst[None], MASS, NITER, NRESTART, RESID)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_uml_deflated.npz")
    np.savez_compressed(path, dims=np.array(DIMS), link_seed=LINK_SEED, src_seed=SRC_SEED, mass=MASS, niter=NITER,
                        nrestart=NRESTART, resid=RESID, nvecs=np.array(ks), iters=np.array(iters), iters_plain=it0,
                        solutions=np.stack(sols), eigval=lam[:48])
    print("wrote", path, os.path.getsize(path), "bytes; iterations", dict(zip(ks, iters)), "undeflated", it0)


if __name__ == "__main__":
    main()

"""Generates tests/golden/ref_l6666_synth.npz from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/build_ref.sh's output):

    python tests/golden/make_golden.py

Inputs are the seeded synthetic HISQ-like fields of milc_qcd_b200.fields on a 6^4 lattice
(BASELINE config 1's volume); outputs come from the reference's own compiled sources
(oracle/_ref/libmilcref.so = generic_ks/dslash_fn_dblstore.c, d_congrad5_fn_milc.c,
ks_multicg_offset.c built with its default flags).  The .npz is committed so the parity
checks also run where /root/reference does not exist (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from milc_qcd_b200 import fields as F  # noqa: E402
from oracle.pyoracle import MilcRef, EVEN, EVENANDODD  # noqa: E402


def main():
    dims = (6, 6, 6, 6)
    mass = 0.05
    fat, lng = F.make_links(dims, seed=1234)
    src = F.make_source(dims, seed=1235, parity=EVENANDODD)
    cg_src = F.make_source(dims, seed=5678, parity=EVEN)
    ref = MilcRef(dims)
    ref.set_links(fat, lng)
    dslash = ref.dslash(src, EVENANDODD)
    cg_niter, cg_nrestart, cg_resid = 500, 5, 1e-10
    x = np.zeros_like(cg_src)
    it, q = ref.congrad(cg_src, x, mass, EVEN, cg_niter, cg_nrestart, cg_resid)
    offsets = np.roll(F.rhmc_offsets(11, mass), 2)
    ms_resid = 1e-8
    ms_it, psim, mq = ref.multicg(cg_src, offsets, EVEN, 2000, 1, ms_resid)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_l6666_synth.npz")
    np.savez(out, dims=np.array(dims), mass=mass, fat=fat, lng=lng, src=src, dslash=dslash,
             cg_src=cg_src, cg_x=x, cg_iters=it, cg_niter=cg_niter, cg_nrestart=cg_nrestart,
             cg_resid=cg_resid, cg_final_rsq=q["final_rsq"], cg_final_restart=q["final_restart"],
             ms_offsets=offsets, ms_resid=ms_resid, ms_iters=ms_it, ms_psim=psim[:, :648].copy(),
             ms_final_rsq=mq[0]["final_rsq"])
    print("wrote", out, os.path.getsize(out), "bytes; cg iters", it, "multicg iters", ms_it)


if __name__ == "__main__":
    main()

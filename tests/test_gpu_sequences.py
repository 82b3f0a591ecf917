"""GPU parity tests of the device-resident solve sequences: the
UML propagator solve (generic_ks/mat_invert.c:328-402,409-475) and the multi-shift solve followed
by the rational-function sum / the other-parity fill of its RHMC callers
(ks_imp_rhmc/ks_ratinv.c:121-138, update_h_rhmc.c:75-86).  The oracle is the reference's sequence
composed from the pinned oracle primitives (dslash, single-mass CG, multi-shift CG)."""
import numpy as np
import pytest

from conftest import fields_for

pytestmark = pytest.mark.gpu

EVEN, ODD, EVENANDODD = 2, 1, 3


@pytest.fixture(scope="module")
def api():
    from milc_qcd_b200 import api
    yield api
    api.finalize()


def uml_oracle(o, dims, fat, lng, src, guess, mass, niter, nrestart, resid):
    """mat_invert_uml_field, generic_ks/mat_invert.c:328-402 (this composition is pinned on the
    reference's compiled function in tests/test_oracle.py)."""
    h = src.shape[0] // 2
    tmp = -o.dslash(dims, fat, lng, src, EVENANDODD) + 2 * mass * src        # M^+ src
    dst = guess.copy()
    it_e, q_e = o.congrad(dims, fat, lng, tmp, dst, mass, EVEN, niter, nrestart, resid)
    ttt = o.dslash(dims, fat, lng, dst, ODD)
    dst[h:] = (src[h:] - ttt[h:]) / (2 * mass)
    it_o, q_o = o.congrad(dims, fat, lng, tmp, dst, mass, ODD, niter, nrestart, resid)
    return dst, it_e, it_o


@pytest.mark.parametrize("dims,nsrc,mixed", [((8, 8, 8, 8), 1, 0), ((8, 12, 6, 10), 3, 0), ((8, 8, 8, 8), 3, 1)])
def test_uml_sequence_matches_reference_sequence(api, oracle, dims, nsrc, mixed):
    from milc_qcd_b200 import fields as F
    fat, lng, _ = fields_for(dims)
    mass, resid = 0.05, 1e-10
    srcs = [F.make_source(dims, seed=900 + 7 * k, parity=EVENANDODD) for k in range(nsrc)]
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    dsts = [np.zeros_like(s) for s in srcs]
    tot, res = ctx.mat_invert_uml(srcs, dsts, mass, 500, 5, resid, mixed_precision=mixed)
    assert tot == sum(e["final_iters"] + o_["final_iters"] for e, o_ in res)
    for k in range(nsrc):
        want, it_e, it_o = uml_oracle(oracle, dims, fat, lng, srcs[k], np.zeros_like(srcs[k]), mass, 500, 5, resid)
        even, odd = res[k]
        assert even["converged"] == 1 and odd["converged"] == 1
        if mixed == 0:
            assert abs(even["final_iters"] - it_e) <= max(2, 0.02 * it_e)
            assert abs(odd["final_iters"] - it_o) <= 2        # the polish: a few iterations at most
        assert np.linalg.norm(dsts[k] - want) <= 10 * resid / (4 * mass * mass) * np.linalg.norm(want)
        # independent check of M dst = src on all sites
        r = oracle.dslash(dims, fat, lng, dsts[k], EVENANDODD) + 2 * mass * dsts[k] - srcs[k]
        assert np.linalg.norm(r) <= 1e-7 * np.linalg.norm(srcs[k])
    # MILC-named entry points
    fn = api.fn_links_t(fat=fat, lng=lng, dims=dims)
    qic = api.quark_invert_control(max=500, nrestart=5, resid=resid, mixed_precision=mixed)
    d1 = np.zeros_like(srcs[0])
    it = api.mat_invert_uml_field(srcs[0], d1, qic, mass, fn)
    assert it == qic.final_iters and qic.converged == 1
    if mixed == 0:   # single-source call vs block call: the same bits
        assert np.array_equal(d1, dsts[0])
    else:
        assert np.linalg.norm(d1 - dsts[0]) <= 1e-7 * np.linalg.norm(dsts[0])
    if nsrc > 1:   # block form, the reference's argument order (mat_invert.c:409-411)
        dn = [np.zeros_like(s) for s in srcs]
        qb = api.quark_invert_control(max=500, nrestart=5, resid=resid, mixed_precision=mixed)
        itb = api.mat_invert_block_uml(srcs, dn, mass, nsrc, qb, fn)
        assert itb == tot and qb.converged == 1 and qb.final_iters == itb
        assert all(np.array_equal(a, b) for a, b in zip(dn, dsts))
    ctx.close()


def test_multishift_rational_sum_and_other_parity_fill(api, oracle):
    from milc_qcd_b200 import fields as F
    dims = (8, 8, 8, 8)
    fat, lng, _ = fields_for(dims)
    src = F.make_source(dims, seed=31, parity=EVEN)
    offsets = np.roll(F.rhmc_offsets(7, 0.05), 3)
    residues = np.array([0.37, 1.0, -0.5, 0.25, 2.0, -0.125, 0.7, 0.01])
    resid = 1e-9
    it_o, p_o, q_o = oracle.multicg(dims, fat, lng, src, offsets, EVEN, 3000, 1, resid)
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    V = src.shape[0]
    h = V // 2
    # plain: the same as b200ks_multicg
    it, psim, dest, res = ctx.multicg_rational(src, offsets, EVEN, 3000, 1, resid)
    ps2 = [np.zeros_like(src) for _ in offsets]
    it2, res2 = ctx.multicg(src, ps2, offsets, EVEN, 3000, 1, resid)
    assert it == it2 and dest is None and all(np.array_equal(a, b) for a, b in zip(psim, ps2))
    assert abs(it - it_o) <= max(2, 0.02 * it_o)
    # rational function only: dest = r0 src + sum r_j psim_j  (ks_rateval)
    it, none, dest, res = ctx.multicg_rational(src, offsets, EVEN, 3000, 1, resid, residues=residues, want_psim=False)
    assert none is None
    want = residues[0] * src + sum(r * p for r, p in zip(residues[1:], p_o))
    assert np.linalg.norm(dest[:h] - want[:h]) <= 1e-7 * np.linalg.norm(want[:h]) and np.all(dest[h:] == 0)
    # both: solutions with the odd sites filled by D (update_h_rhmc.c:82-84) + the sum
    it, psim, dest, res = ctx.multicg_rational(src, offsets, EVEN, 3000, 1, resid, residues=residues, fill_other=True)
    for j in range(len(offsets)):
        assert np.array_equal(psim[j][:h], ps2[j][:h])
        wo = oracle.dslash(dims, fat, lng, p_o[j], ODD)
        assert np.linalg.norm(psim[j][h:] - wo[h:]) <= 1e-6 * np.linalg.norm(wo[h:])
    ctx.close()

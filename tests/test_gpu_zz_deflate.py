"""GPU parity tests of the low-mode deflation: b200ks_eig_set /
b200ks_deflate_dev and the deflated UML sequence through the C ABI, against the CPU oracle
(oracle/ks_oracle.c kso_deflate) and the committed output of the reference's mat_invert_uml_field with
qic->deflate = 1 (tests/golden/ref_uml_deflated.npz).  The kernels' arithmetic is checked on the host in
tests/test_deflate_host.py.  (Written after the round's GPU budget was spent: this file sorts last.)"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EVEN, ODD, EVENANDODD = 2, 1, 3


@pytest.fixture(scope="module")
def api():
    from milc_qcd_b200 import api
    yield api
    api.finalize()


@pytest.fixture(scope="module")
def case(oracle):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_deflate import low_modes
    from milc_qcd_b200 import fields as F
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_uml_deflated.npz"))
    dims = tuple(int(d) for d in g["dims"])
    fat, lng = F.make_links(dims, seed=int(g["link_seed"]))
    lam, ev = low_modes(oracle, dims, fat, lng)
    src = F.make_source(dims, seed=int(g["src_seed"]), parity=EVENANDODD)
    return g, dims, fat, lng, lam, ev, src


def _upload_modes(ctx, ev, k):
    hs = []
    for j in range(k):
        h = ctx.vec_create()
        ctx.vec_upload(h, np.ascontiguousarray(ev[j]), EVENANDODD)
        hs.append(h)
    return hs


def test_deflate_matches_oracle(api, oracle, case):
    g, dims, fat, lng, lam, ev, src = case
    ctx = api.Context(dims)
    k = 16
    hs = _upload_modes(ctx, ev, k)
    ctx.eig_set(hs, lam[:k], use_in_uml=False)
    assert ctx.eig_count() == k
    vs, vd = ctx.vec_create(), ctx.vec_create()
    rng = np.random.default_rng(6)
    guess = rng.standard_normal(src.shape)
    ctx.vec_upload(vs, src, EVENANDODD)
    for parity in (EVEN, ODD):
        ctx.vec_upload(vd, guess, EVENANDODD)
        ctx.deflate_dev(vs, vd, 0.03, parity)
        got = ctx.vec_download(vd, np.zeros_like(src), EVENANDODD)
        want = oracle.deflate(dims, guess.copy(), src, 0.03, ev[:k], lam[:k], parity)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    with pytest.raises(Exception):        # a vector of the set cannot be freed
        ctx.vec_free(hs[0])
    ctx.eig_set([], [])
    assert ctx.eig_count() == 0
    ctx.vec_free(hs[0])
    ctx.close()


def test_deflated_uml_matches_reference_golden(api, oracle, case):
    """Iteration counts (the CG trajectory depends on the trial solution) and solutions of the reference's
    deflated mat_invert_uml_field for 8, 48 and all 384 low modes; with all of them both CGs stop at their
    first true-residual check."""
    g, dims, fat, lng, lam, ev, src = case
    mass, niter, nrestart, resid = float(g["mass"]), int(g["niter"]), int(g["nrestart"]), float(g["resid"])
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    hs = _upload_modes(ctx, ev, len(lam))
    for k, it_ref, want in zip(g["nvecs"], g["iters"], g["solutions"]):
        k = int(k)
        ctx.eig_set(hs[:k], lam[:k], use_in_uml=True)
        dst = np.zeros_like(src)
        tot, res = ctx.mat_invert_uml([src], [dst], mass, niter, nrestart, resid)
        even, odd = res[0]
        assert even["converged"] == 1 and odd["converged"] == 1
        assert abs(tot - int(it_ref)) <= max(2, 0.02 * int(it_ref)), (k, tot, it_ref)
        assert np.linalg.norm(dst - want) <= 1e-8 * np.linalg.norm(want), k
    assert tot <= 4                          # reference: 2 (one true-residual check per parity)
    # without the set: the plain sequence again
    ctx.eig_set([], [])
    dst = np.zeros_like(src)
    tot, _ = ctx.mat_invert_uml([src], [dst], mass, niter, nrestart, resid)
    assert tot > 2 * int(g["iters"][0])      # reference: 1226 against 438 with 8 modes
    ctx.close()

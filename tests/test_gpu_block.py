"""GPU parity tests of the multi-right-hand-side (block) path: the K-wide stencil and the block
CG behind ks_congrad_block_parity_gpu / qudaInvertMsrc.  The reference's block solver is a loop
of single solves (generic_ks/d_congrad5_fn_milc.c:409-417), so the oracle for K sources is the
oracle's single solve applied K times; on top of that the block path must reproduce this
library's own single-source kernels bit for bit (same arithmetic per right-hand side)."""
import numpy as np
import pytest

from conftest import fields_for

pytestmark = pytest.mark.gpu

EVEN, ODD, EVENANDODD = 2, 1, 3
DSLASH_TOL = 1e-13


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def api():
    from milc_qcd_b200 import api
    yield api
    api.finalize()


def _sources(dims, n, parity, seed0=700):
    from milc_qcd_b200 import fields as F
    return [F.make_source(dims, seed=seed0 + 13 * k, parity=parity) for k in range(n)]


@pytest.mark.parametrize("dims", [(8, 8, 8, 8), (8, 12, 6, 10), (4, 4, 4, 4)])
@pytest.mark.parametrize("long_recon", [18, 14])
def test_block_dslash_matches_oracle_and_single_kernel(api, oracle, dims, long_recon):
    fat, lng, _ = fields_for(dims)
    ctx = api.Context(dims)
    ctx.load_links(fat, lng, long_recon)
    srcs = _sources(dims, 4, EVENANDODD)
    vs = [ctx.vec_create() for _ in range(4)]
    vd = [ctx.vec_create() for _ in range(4)]
    v1 = ctx.vec_create()
    for k in range(4):
        ctx.vec_upload(vs[k], srcs[k])
    want = [oracle.dslash(dims, fat, lng, s, EVENANDODD) for s in srcs]
    for prec, tol in ((2, DSLASH_TOL), (1, 2e-6)):
        for nrhs in (1, 2, 3, 4):
            for parity in (EVEN, ODD, EVENANDODD):
                for k in range(nrhs):
                    ctx.vec_zero(vd[k])
                ctx.dslash_block_dev(vs[:nrhs], vd[:nrhs], parity, prec)
                V = srcs[0].shape[0]
                sl = slice(0, V // 2) if parity == EVEN else slice(V // 2, V) if parity == ODD else slice(0, V)
                for k in range(nrhs):
                    got = np.zeros_like(srcs[k])
                    ctx.vec_download(vd[k], got)
                    assert rel_err(got[sl], want[k][sl]) <= tol, (prec, nrhs, parity, k)
                    mask = np.ones(V, bool)
                    mask[sl] = False
                    assert np.all(got[mask] == 0)          # only `parity` sites are written
                    # the same bits as the single-source stencil
                    ctx.vec_zero(v1)
                    ctx.dslash_dev(vs[k], v1, parity, prec)
                    one = np.zeros_like(srcs[k])
                    ctx.vec_download(v1, one)
                    assert np.array_equal(got, one), (prec, nrhs, parity, k)
    ctx.close()


@pytest.mark.parametrize("dims,parity,nsrc", [((8, 8, 8, 8), EVEN, 3), ((8, 12, 6, 10), ODD, 4), ((6, 6, 6, 6), EVEN, 2)])
def test_block_congrad_reproduces_single_solves(api, oracle, dims, parity, nsrc):
    """mixed_precision 0: per right-hand side the block solve is the single solve -- same
    iteration counts, restarts, residuals and solution bits -- and therefore matches the oracle
    exactly as the single solve does."""
    fat, lng, _ = fields_for(dims)
    srcs = _sources(dims, nsrc, parity)
    srcs[1] = 37.0 * srcs[1]          # different norms: per-source residual targets
    mass, resid = 0.05, 1e-10
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    xs = [np.zeros_like(s) for s in srcs]
    tot, res = ctx.congrad_block(srcs, xs, mass, parity, 500, 5, resid)
    assert tot == sum(r["final_iters"] for r in res)
    V = srcs[0].shape[0]
    sl = slice(0, V // 2) if parity == EVEN else slice(V // 2, V)
    for k in range(nsrc):
        x1 = np.zeros_like(srcs[k])
        it1, r1 = ctx.congrad(srcs[k], x1, mass, parity, 500, 5, resid)
        assert res[k]["final_iters"] == it1 and res[k]["final_restart"] == r1["final_restart"]
        assert res[k]["converged"] == 1 and res[k]["final_rsq"] == r1["final_rsq"]
        assert np.array_equal(xs[k], x1)
        xo = np.zeros_like(srcs[k])
        ito, qo = oracle.congrad(dims, fat, lng, srcs[k], xo, mass, parity, 500, 5, resid)
        assert abs(res[k]["final_iters"] - ito) <= max(2, 0.02 * ito)
        assert np.linalg.norm(xs[k] - xo) <= 10 * resid / (4 * mass * mass) * np.linalg.norm(xo)
        mask = np.ones(V, bool)
        mask[sl] = False
        assert np.all(xs[k][mask] == 0)
    ctx.close()


def test_block_congrad_edge_cases(api, oracle):
    """Zero source inside a block, more sources than one pass holds (4 + 2), initial guesses, an
    iteration cap that only some sources hit, and the MILC-named entry point."""
    dims = (6, 6, 6, 6)
    fat, lng, _ = fields_for(dims)
    mass = 0.05
    srcs = _sources(dims, 6, EVEN)
    srcs[2] = np.zeros_like(srcs[2])
    V = srcs[0].shape[0]
    fn = api.fn_links_t(fat=fat, lng=lng, dims=dims)
    xs = [np.ones_like(s) for s in srcs]
    for x in xs:
        x[:V // 2] = 0
    qic = api.quark_invert_control(max=500, nrestart=5, parity=EVEN, resid=1e-9)
    tot = api.ks_congrad_block_parity_gpu(6, srcs, xs, qic, mass, fn)
    assert qic.converged == 1 and qic.final_iters == tot and qic.final_rsq < 1e-18
    singles = 0
    for k in range(6):
        assert np.all(xs[k][V // 2:] == 1)           # odd half untouched
        x1 = np.zeros_like(srcs[k])
        q1 = api.quark_invert_control(max=500, nrestart=5, parity=EVEN, resid=1e-9)
        singles += api.ks_congrad_parity_gpu(srcs[k], x1, q1, mass, fn)
        assert np.array_equal(xs[k][:V // 2], x1[:V // 2])
    assert tot == singles
    assert np.all(xs[2][:V // 2] == 0)               # zero source -> zero solution, no iterations
    # converged solutions as initial guesses: one true-residual evaluation each (5 live sources)
    q2 = api.quark_invert_control(max=500, nrestart=5, parity=EVEN, resid=1e-8)
    assert api.ks_congrad_block_parity_gpu(6, srcs, xs, q2, mass, fn) == 5
    # iteration cap: sources 0 and 1 are hard (need > 21 iterations), source 2 starts converged
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    guess = [np.zeros_like(srcs[0]), np.zeros_like(srcs[1]), xs[3].copy()]
    tot, res = ctx.congrad_block([srcs[0], srcs[1], srcs[3]], guess, mass, EVEN, 7, 3, 1e-8)
    for k, s in enumerate((srcs[0], srcs[1])):
        xo = np.zeros_like(s)
        ito, qo = oracle.congrad(dims, fat, lng, s, xo, mass, EVEN, 7, 3, 1e-8)
        assert (res[k]["final_iters"], res[k]["converged"], res[k]["final_restart"]) == (ito, qo["converged"], qo["final_restart"])
        assert np.abs(guess[k] - xo).max() <= 1e-9 * np.abs(xo).max()
    assert res[2]["final_iters"] == 1 and res[2]["converged"] == 1
    # no sources at all
    assert ctx.congrad_block([], [], mass, EVEN, 7, 3, 1e-8)[0] == 0
    ctx.close()


@pytest.mark.parametrize("dims,parity,nsrc", [((8, 8, 8, 8), EVEN, 3), ((8, 12, 6, 10), ODD, 4)])
def test_block_congrad_mixed_precision(api, oracle, dims, parity, nsrc):
    """mixed_precision != 0: single-precision Krylov vectors, K at a time, joint reliable updates;
    every solution meets the double-precision true residual and agrees with the oracle."""
    fat, lng, _ = fields_for(dims)
    srcs = _sources(dims, nsrc, parity)
    srcs[0] = 1e-3 * srcs[0]
    mass, resid = 0.05, 1e-10
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    xs = [np.zeros_like(s) for s in srcs]
    tot, res = ctx.congrad_block(srcs, xs, mass, parity, 500, 5, resid, mixed_precision=1)
    V = srcs[0].shape[0]
    sl = slice(0, V // 2) if parity == EVEN else slice(V // 2, V)
    op = EVEN if parity == ODD else ODD
    for k in range(nsrc):
        assert res[k]["converged"] == 1 and res[k]["final_rsq"] < resid ** 2
        xo = np.zeros_like(srcs[k])
        ito, qo = oracle.congrad(dims, fat, lng, srcs[k], xo, mass, parity, 500, 5, resid)
        assert res[k]["final_iters"] <= 1.5 * ito + 20, (k, res[k]["final_iters"], ito)
        t = oracle.dslash(dims, fat, lng, oracle.dslash(dims, fat, lng, xs[k], op), parity)
        r = srcs[k][sl] - (4 * mass * mass * xs[k][sl] - t[sl])
        assert np.linalg.norm(r) <= 2 * resid * np.linalg.norm(srcs[k][sl])
        assert np.linalg.norm(xs[k] - xo) <= 10 * resid / (4 * mass * mass) * np.linalg.norm(xo)
    ctx.close()


@pytest.mark.parametrize("force,dims", [("t", (8, 6, 8, 12)), ("zt", (8, 6, 8, 12)), ("zt", (4, 4, 6, 6)), ("zt", (8, 8, 4, 4))])
def test_block_and_sequence_calls_on_a_partitioned_context(api, oracle, monkeypatch, force, dims):
    """A partitioned context (here one GPU as its own neighbour: push kernels, ghost buffers, arrival flags, split
    reductions all run) applies the K-wide stencil too (round 2): one exchange carries the halos of all K inputs.  The
    block stencil gives the single-source stencil's bits, the pure-double block CG the single solves' bits, the mixed
    block CG meets the double true residual; the UML sequence of several sources goes through it.  The link
    construction still refuses a partitioned context with a clear error."""
    from milc_qcd_b200 import fields as F
    monkeypatch.setenv("B200KS_FORCE_PARTITION", force)
    fat, lng, _ = fields_for(dims)
    ctx = api.Context(dims, grid=(1, 1, 1, 1), rank=0, nranks=1)
    assert ctx.halo_mode() == 2
    ctx.load_links(fat, lng)
    V = int(np.prod(dims))
    # K-wide stencil on the partitioned lattice: oracle, and the single-source kernel's bits
    full = _sources(dims, 4, EVENANDODD, seed0=900)
    vs = [ctx.vec_create() for _ in range(4)]
    vd = [ctx.vec_create() for _ in range(4)]
    v1 = ctx.vec_create()
    for k in range(4):
        ctx.vec_upload(vs[k], full[k])
    want = [oracle.dslash(dims, fat, lng, f, EVENANDODD) for f in full]
    for prec, tol in ((2, DSLASH_TOL), (1, 2e-6)):
        for nrhs in (2, 3, 4):
            for parity in (EVEN, ODD):
                for k in range(nrhs):
                    ctx.vec_zero(vd[k])
                ctx.dslash_block_dev(vs[:nrhs], vd[:nrhs], parity, prec)
                sl = slice(0, V // 2) if parity == EVEN else slice(V // 2, V)
                for k in range(nrhs):
                    got = np.zeros_like(full[k])
                    ctx.vec_download(vd[k], got)
                    assert rel_err(got[sl], want[k][sl]) <= tol, (prec, nrhs, parity, k)
                    ctx.vec_zero(v1)
                    ctx.dslash_dev(vs[k], v1, parity, prec)
                    one = np.zeros_like(full[k])
                    ctx.vec_download(v1, one)
                    assert np.array_equal(got, one), (prec, nrhs, parity, k)
    # block CG
    srcs = _sources(dims, 3, EVEN)
    srcs[1] = 37.0 * srcs[1]
    xs = [np.zeros_like(s) for s in srcs]
    tot, res = ctx.congrad_block(srcs, xs, 0.05, EVEN, 500, 5, 1e-9)
    for k in range(3):
        x1 = np.zeros_like(srcs[k])
        it1, r1 = ctx.congrad(srcs[k], x1, 0.05, EVEN, 500, 5, 1e-9)
        assert res[k]["final_iters"] == it1 and res[k]["final_rsq"] == r1["final_rsq"] and res[k]["converged"] == 1
        assert np.array_equal(xs[k], x1)
        xo = np.zeros_like(srcs[k])
        ito, qo = oracle.congrad(dims, fat, lng, srcs[k], xo, 0.05, EVEN, 500, 5, 1e-9)
        assert abs(res[k]["final_iters"] - ito) <= max(2, 0.02 * ito)
        assert np.linalg.norm(xs[k] - xo) <= 1e-7 * np.linalg.norm(xo)
    xm = [np.zeros_like(s) for s in srcs]
    tot, resm = ctx.congrad_block(srcs, xm, 0.05, EVEN, 500, 5, 1e-9, mixed_precision=1)
    for k in range(3):
        assert resm[k]["converged"] == 1 and resm[k]["final_rsq"] < 1e-18
        assert np.linalg.norm(xm[k] - xs[k]) <= 1e-7 * np.linalg.norm(xs[k])
    # the resident UML sequence with several sources (block solver underneath)
    fulls = [F.make_source(dims, seed=77 + k, parity=EVENANDODD) for k in range(2)]
    dsts = [np.zeros_like(f) for f in fulls]
    it, r = ctx.mat_invert_uml(fulls, dsts, 0.05, 500, 5, 1e-9)
    for k in range(2):
        resid = oracle.dslash(dims, fat, lng, dsts[k], EVENANDODD) + 0.1 * dsts[k] - fulls[k]
        assert np.linalg.norm(resid) <= 1e-6 * np.linalg.norm(fulls[k])
    with pytest.raises(Exception, match="single-GPU"):
        ctx.hisq_links(F.make_thin_links(dims, seed=5))
    ctx.close()


def test_block_argument_errors(api):
    dims = (4, 4, 4, 4)
    fat, lng, src = fields_for(dims)
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    v = [ctx.vec_create() for _ in range(6)]
    with pytest.raises(Exception, match="distinct|different"):
        ctx.congrad_block_dev([v[0], v[1]], [v[2], v[2]], 0.05, EVEN, 10, 1, 1e-6)      # same solution field twice
    with pytest.raises(Exception, match="different"):
        ctx.congrad_block_dev([v[0]], [v[0]], 0.05, EVEN, 10, 1, 1e-6)
    with pytest.raises(Exception, match="parity"):
        ctx.congrad_block([src], [np.zeros_like(src)], 0.05, EVENANDODD, 10, 1, 1e-6)
    with pytest.raises(Exception, match="1..4"):
        ctx.dslash_block_dev(v[:5], v[:5], EVEN, 2)
    with pytest.raises(Exception, match="2m"):
        ctx.mat_invert_uml([src], [np.zeros_like(src)], 0.0, 10, 1, 1e-6)
    with pytest.raises(Exception, match="path coefficients"):
        ctx.ks_links(np.zeros((src.shape[0], 4, 3, 3, 2)), (1.0, 2.0))
    ctx.close()

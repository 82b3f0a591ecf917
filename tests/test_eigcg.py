"""eigCG / incremental eigCG (generic_ks/inc_eigcg.c).

CPU: oracle/eigcg_oracle.py (numpy restatement) is pinned live on the reference's own inc_eigcg.c compiled into
oracle/_ref (when that build found a LAPACK) and on the committed golden tests/golden/ref_eigcg.npz that build
produced (tests/golden/make_golden_eigcg.py).  GPU: the library's eigCG through the C ABI against the same golden
and the oracle.  What can be compared: iteration counts (the CG part is an ordinary CG; +-2 from summation order),
converged solutions, Ritz VALUES (the low ones to 1e-6: they converge as vectors accumulate; single-solve values are
rough by construction and move at the 1e-3 level with one iteration more or less) and the SPAN of the vectors."""
import os

import numpy as np
import pytest

from conftest import ROOT

EVEN = 2
GOLD = os.path.join(ROOT, "tests", "golden", "ref_eigcg.npz")


def _setup():
    from milc_qcd_b200 import fields as F
    g = np.load(GOLD)
    dims = tuple(int(x) for x in g["dims"])
    fat, lng = F.make_links(dims, seed=1234)
    return g, dims, fat, lng, F


def _span_defect(vecs, ref_vecs):
    """largest distance of a reference vector from the span of `vecs` (complex vectors as rows)."""
    Q, _ = np.linalg.qr(np.array(vecs).T)
    R = np.array(ref_vecs).T
    return float(np.abs(R - Q @ (Q.conj().T @ R)).max() / np.abs(R).max())


def test_eigcg_oracle_reproduces_the_reference_golden(oracle):
    from oracle.eigcg_oracle import EigCGOracle
    g, dims, fat, lng, F = _setup()
    V = int(np.prod(dims))
    eo = EigCGOracle(oracle, dims, fat, lng)
    eo.inc_init(int(g["m"]), int(g["nvecs"]), int(g["nmax"]))
    for s in range(len(g["iters"])):
        b = F.make_source(dims, seed=2000 + s, parity=EVEN)
        xc = np.zeros(V // 2 * 3, complex)
        it, q = eo.inc_eigcg(eo.to_c(b, EVEN), xc, float(g["mass"]), EVEN, 2000, 5, float(g["resid"]))
        assert abs(it - int(g["iters"][s])) <= 4, (s, it, g["iters"][s])
        assert eo.p["Nvecs_curr"] == int(g["ncurr"][s]) and q["converged"] == 1
        ref = (g["sols"][s][..., 0] + 1j * g["sols"][s][..., 1]).reshape(-1)
        assert np.linalg.norm(xc - ref) <= 1e-8 * np.linalg.norm(ref)
    w, vecs = eo.pairs()
    assert np.abs(w[:6] - g["eigval"][:6]).max() <= 1e-6 * np.abs(g["eigval"][:6]).max()
    assert np.all(np.abs(w - g["eigval"]) <= 0.15 * np.abs(g["eigval"]))
    gv = (g["eigvec_even"][..., 0] + 1j * g["eigvec_even"][..., 1]).reshape(len(w), -1)
    assert _span_defect(vecs, gv[:6]) <= 2e-3
    # incremental eigCG pays: the last solve needs 15 % fewer iterations than the first
    assert g["iters"][-1] < 0.85 * g["iters"][0]


def test_eigcg_oracle_matches_the_compiled_reference_live(oracle):
    from oracle import pyoracle
    from oracle.eigcg_oracle import EigCGOracle
    if not pyoracle.ref_available(""):
        pytest.skip("oracle/_ref not built")
    g, dims, fat, lng, F = _setup()
    ref = pyoracle.MilcRef(dims, "")
    if not ref.has_eigcg:
        pytest.skip("oracle/_ref was built without inc_eigcg.c (no LAPACK at build time)")
    ref.set_links(fat, lng)
    V = int(np.prod(dims))
    eo = EigCGOracle(oracle, dims, fat, lng)
    b = F.make_source(dims, seed=77, parity=EVEN)
    for (m, nv, cap) in ((40, 6, 150), (16, 4, 97)):
        # a fixed number of iterations (unreachable target, one restart interval): the same Krylov space in both
        x = np.zeros_like(b)
        it, val, vec, q = ref.eigcg(b, x, 0.05, EVEN, cap, 1, 1e-30, m, nv)
        xc = np.zeros(V // 2 * 3, complex)
        work = [None] * m
        it2, val2, q2 = eo.eigcg(eo.to_c(b, EVEN), xc, 0.05, EVEN, cap, 1, 1e-30, m, nv, work)
        assert it == it2 == cap + 1
        assert np.abs(val - val2).max() <= 2e-4 * np.abs(val).max()
        for j in range(nv):
            assert abs(np.vdot(eo.to_c(vec[j], EVEN), work[j])) >= 1 - 1e-4
        assert np.linalg.norm(eo.to_c(x, EVEN) - xc) <= 1e-4 * np.linalg.norm(xc)


@pytest.mark.gpu
def test_inc_eigcg_on_the_gpu_matches_the_reference_golden(oracle):
    from milc_qcd_b200 import api
    g, dims, fat, lng, F = _setup()
    V = int(np.prod(dims))
    ctx = api.Context(dims)
    ctx.load_links(fat, lng)
    ctx.eigcg_init(int(g["m"]), int(g["nvecs"]), int(g["nmax"]))
    plain = []
    for s in range(len(g["iters"])):
        b = F.make_source(dims, seed=2000 + s, parity=EVEN)
        x = np.zeros_like(b)
        it, res = ctx.inc_eigcg(b, x, float(g["mass"]), EVEN, 2000, 5, float(g["resid"]))
        assert res["converged"] == 1 and res["final_rsq"] < float(g["resid"]) ** 2
        assert abs(it - int(g["iters"][s])) <= max(4, 0.02 * int(g["iters"][s])), (s, it, g["iters"][s])
        assert ctx.eigcg_count() == int(g["ncurr"][s])
        assert np.linalg.norm(x[: V // 2] - g["sols"][s]) <= 1e-8 * np.linalg.norm(g["sols"][s])
        assert np.all(x[V // 2:] == 0)
        x0 = np.zeros_like(b)
        plain.append(ctx.congrad(b, x0, float(g["mass"]), EVEN, 2000, 5, float(g["resid"]))[0])
    assert it < 0.9 * plain[-1]          # the accumulated vectors do deflate the solve
    w = ctx.eigcg_pairs()
    assert len(w) == int(g["ncurr"][-1]) and np.all(np.diff(w) >= 0)
    assert np.abs(w[:6] - g["eigval"][:6]).max() <= 1e-6 * np.abs(g["eigval"][:6]).max()
    assert np.all(np.abs(w - g["eigval"]) <= 0.15 * np.abs(g["eigval"]))
    vecs = []
    for j in range(len(w)):
        v = ctx.eigcg_vec(j)
        vecs.append((v[: V // 2, :, 0] + 1j * v[: V // 2, :, 1]).reshape(-1))
    G = np.array(vecs)
    assert np.abs(G.conj() @ G.T - np.eye(len(w))).max() <= 1e-10          # orthonormal
    gv = (g["eigvec_even"][..., 0] + 1j * g["eigvec_even"][..., 1]).reshape(len(w), -1)
    assert _span_defect(vecs, gv[:6]) <= 2e-3
    # each is an approximate eigenvector of -D_eo D_oe with its Ritz value
    for j in range(4):
        f = np.zeros((V, 3, 2))
        f[: V // 2, :, 0], f[: V // 2, :, 1] = vecs[j].real.reshape(-1, 3), vecs[j].imag.reshape(-1, 3)
        t = oracle.dslash(dims, fat, lng, oracle.dslash(dims, fat, lng, f, 1), EVEN)
        Av = -(t[: V // 2, :, 0] + 1j * t[: V // 2, :, 1]).reshape(-1)
        assert abs(np.vdot(vecs[j], Av).real - w[j]) <= 1e-10
        assert np.linalg.norm(Av - w[j] * vecs[j]) <= 2e-2 * np.sqrt(abs(w[j]))
    ctx.close()


@pytest.mark.gpu
def test_eigcg_refuses_bad_parameters():
    from milc_qcd_b200 import api, _lib
    ctx = api.Context((4, 4, 4, 4))
    with pytest.raises(_lib.B200KSError):
        ctx.eigcg_init(10, 5, 20)        # 2 Nvecs must be < m
    with pytest.raises(_lib.B200KSError):
        ctx.inc_eigcg(np.zeros((256, 3, 2)), np.zeros((256, 3, 2)), 0.1, EVEN, 10, 1, 1e-6)   # no init, no links
    ctx.close()

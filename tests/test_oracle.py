"""CPU tests that PIN the oracle (oracle/ks_oracle.c):
  * against the committed golden vectors produced by the reference's own compiled CPU path
    (tests/golden/ref_l6666_synth.npz, generator tests/golden/make_golden.py);
  * live against oracle/_ref/libmilcref.so (the reference's sources compiled by
    oracle/build_ref.sh) whenever that library is present.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, fields_for

EVEN, ODD, EVENANDODD = 2, 1, 3
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_l6666_synth.npz")


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_node_index_is_milc_layout(oracle):
    from milc_qcd_b200 import fields as F
    dims = (4, 6, 2, 8)
    perm = F.lex_to_milc(dims)
    nx, ny, nz, nt = dims
    assert sorted(perm) == list(range(nx * ny * nz * nt))
    for (x, y, z, t) in [(0, 0, 0, 0), (1, 0, 0, 0), (3, 5, 1, 7), (2, 3, 1, 4)]:
        lex = x + nx * (y + ny * (z + nz * t))
        assert oracle.node_index(dims, x, y, z, t) == perm[lex]
    V = nx * ny * nz * nt
    assert oracle.node_index(dims, 0, 0, 0, 0) == 0 and oracle.node_index(dims, 1, 0, 0, 0) == V // 2


def test_oracle_dslash_golden(oracle):
    g = np.load(GOLDEN)
    dims = tuple(int(d) for d in g["dims"])
    got = oracle.dslash(dims, g["fat"], g["lng"], g["src"], EVENANDODD)
    assert rel_err(got, g["dslash"]) < 1e-14


def test_oracle_dslash_only_writes_parity(oracle):
    dims = (4, 4, 4, 4)
    fat, lng, src = fields_for(dims)
    full = oracle.dslash(dims, fat, lng, src, EVENANDODD)
    V = src.shape[0]
    for par, sl in ((EVEN, slice(0, V // 2)), (ODD, slice(V // 2, V))):
        dest = np.full_like(src, 5.0)
        oracle.dslash(dims, fat, lng, src, par, dest)
        assert np.array_equal(dest[sl], full[sl])
        mask = np.ones(V, bool)
        mask[sl] = False
        assert np.all(dest[mask] == 5.0)
    buf = src.copy()
    oracle.dslash(dims, fat, lng, buf, EVEN, buf)  # in place, one parity
    assert np.array_equal(buf[:V // 2], full[:V // 2])


def test_oracle_congrad_golden(oracle):
    g = np.load(GOLDEN)
    dims = tuple(int(d) for d in g["dims"])
    x = np.zeros_like(g["cg_src"])
    it, q = oracle.congrad(dims, g["fat"], g["lng"], g["cg_src"], x, float(g["mass"]), EVEN,
                           int(g["cg_niter"]), int(g["cg_nrestart"]), float(g["cg_resid"]))
    assert abs(it - int(g["cg_iters"])) <= 2
    assert q["converged"] == 1 and q["final_restart"] == int(g["cg_final_restart"])
    assert np.linalg.norm(x - g["cg_x"]) <= 1e-9 * np.linalg.norm(g["cg_x"])


def test_oracle_multicg_golden(oracle):
    g = np.load(GOLDEN)
    dims = tuple(int(d) for d in g["dims"])
    it, psim, q = oracle.multicg(dims, g["fat"], g["lng"], g["cg_src"], g["ms_offsets"], EVEN, 2000, 1,
                                 float(g["ms_resid"]))
    assert abs(it - int(g["ms_iters"])) <= 1
    V = g["cg_src"].shape[0]
    assert np.linalg.norm(psim[:, :V // 2] - g["ms_psim"]) <= 1e-8 * np.linalg.norm(g["ms_psim"])
    assert all(qq["converged"] == 1 for qq in q)


def test_oracle_zero_source_and_iteration_cap(oracle):
    dims = (4, 4, 4, 4)
    fat, lng, src = fields_for(dims)
    x = np.ones_like(src)
    it, q = oracle.congrad(dims, fat, lng, np.zeros_like(src), x, 0.05, EVEN, 10, 2, 1e-8)
    V = src.shape[0]
    assert it == 0 and np.all(x[:V // 2] == 0) and np.all(x[V // 2:] == 1)
    x = np.zeros_like(src)
    src_e = src.copy()
    src_e[V // 2:] = 0
    it, q = oracle.congrad(dims, fat, lng, src_e, x, 0.01, EVEN, 5, 2, 1e-14)
    # restart(1) + 4 iterations, restart(6) + 4 iterations, restart(11) >= max_cg: the reference counts
    # every true-residual evaluation as an iteration (d_congrad5_fn_milc.c:223)
    assert it == 11 and q["converged"] == 0


_LIVE = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from milc_qcd_b200 import fields as F
from oracle.pyoracle import Oracle, MilcRef, EVEN, ODD, EVENANDODD
dims = %(dims)r
fat, lng = F.make_links(dims, seed=77)
src = F.make_source(dims, seed=78, parity=EVENANDODD)
o, r = Oracle(), MilcRef(dims)
for c in [(0,0,0,0),(1,0,0,0),(3,1,2,5),(dims[0]-1,dims[1]-1,dims[2]-1,dims[3]-1)]:
    assert o.node_index(dims,*c) == r.node_index(*c)
r.set_links(fat, lng)
for par in (EVEN, ODD, EVENANDODD):
    a, b = o.dslash(dims, fat, lng, src, par), r.dslash(src, par)
    assert np.abs(a-b).max() <= 1e-14*np.abs(b).max(), par
se = F.make_source(dims, seed=79, parity=ODD)
x0, x1 = np.zeros_like(se), np.zeros_like(se)
i0, q0 = o.congrad(dims, fat, lng, se, x0, 0.1, ODD, 300, 5, 1e-9)
i1, q1 = r.congrad(se, x1, 0.1, ODD, 300, 5, 1e-9)
assert abs(i0-i1) <= 2 and q0['converged'] == q1['converged'] == 1
assert q0['final_restart'] == q1['final_restart']
assert np.linalg.norm(x0-x1) <= 1e-8*np.linalg.norm(x1)
offs = F.rhmc_offsets(9, 0.1)[::-1].copy()
i0, p0, _ = o.multicg(dims, fat, lng, se, offs, ODD, 3000, 1, 1e-7)
i1, p1, _ = r.multicg(se, offs, ODD, 3000, 1, 1e-7)
assert i0 == i1 or abs(i0-i1) <= 1
assert np.linalg.norm(p0-p1) <= 1e-8*np.linalg.norm(p1)
print('LIVE-OK')
"""


@pytest.mark.parametrize("dims", [(8, 12, 6, 10), (4, 4, 4, 8)])
def test_oracle_matches_compiled_reference_live(dims):
    from oracle.pyoracle import ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libmilcref.so not built (needs /root/reference)")
    # MILC keeps its geometry in process globals: one subprocess per lattice size
    out = subprocess.run([sys.executable, "-c", _LIVE % dict(root=ROOT, dims=dims)], capture_output=True,
                         text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- HISQ link construction (oracle/ks_links_oracle.c) -------------------------------------------
GOLDEN_LINKS = os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_links.npz")


@pytest.fixture(scope="module")
def links_oracle():
    from oracle.pyoracle import LinksOracle
    return LinksOracle()


@pytest.mark.parametrize("tag,tol_w", [("smooth", 1e-13), ("rough", 1e-10)])
def test_links_oracle_golden(links_oracle, tag, tol_w):
    """The restated chain U -> V -> W -> (fat, long) against the reference's own
    create_hisq_links_milc output (tests/golden/make_golden_links.py), with the reference's
    coefficients; the rough field exercises the SVD branch."""
    from milc_qcd_b200 import fields as F
    g = np.load(GOLDEN_LINKS)
    dims = tuple(int(d) for d in g["dims"])
    U = F.make_thin_links(dims, seed=4321, spread=float(g[tag + "_spread"]))
    assert np.allclose(g["coeffs"][0], links_oracle.FAT7) and np.allclose(g["coeffs"][1], links_oracle.ASQTAD_LIKE)
    o = links_oracle.hisq_links(dims, U)
    assert o["nsvd"] == int(g[tag + "_nsvd"])
    assert rel_err(o["V"], g[tag + "_V"]) < 1e-14
    assert rel_err(o["W"], g[tag + "_W"]) < tol_w
    assert rel_err(o["fat"], g[tag + "_fat"]) < tol_w
    assert rel_err(o["lng"], g[tag + "_lng"]) < tol_w
    # the pieces on their own
    fat, lng = links_oracle.smear(dims, g[tag + "_W"], links_oracle.ASQTAD_LIKE)
    assert rel_err(fat, g[tag + "_fat"]) < 1e-14 and rel_err(lng, g[tag + "_lng"]) < 1e-14
    W, n = links_oracle.unitarize(g[tag + "_V"])
    assert n == int(g[tag + "_nsvd"]) and rel_err(W, g[tag + "_W"]) < tol_w
    Wc = W[..., 0] + 1j * W[..., 1]
    assert np.abs(Wc @ np.conj(np.swapaxes(Wc, -1, -2)) - np.eye(3)).max() < 1e-10


def test_links_oracle_svd_branch_and_degenerate_inputs(links_oracle):
    """Analytic and SVD branches agree where both are valid; unit matrices (degenerate
    eigenvalues, S = 0) and scaled unitary matrices project onto themselves."""
    rng = np.random.default_rng(5)
    V = rng.standard_normal((200, 3, 3, 2))
    Wa, na = links_oracle.unitarize(V, allow_svd=False)
    Ws, ns = links_oracle.unitarize(V, allow_svd=True, svd_rel=0.0, svd_abs=1e300)   # force the SVD branch
    assert na == 0 and ns == 200
    assert np.abs(Wa - Ws).max() < 1e-9
    eye = np.zeros((3, 3, 3, 2))
    for k, s in enumerate((1.0, 2.5, -0.3)):
        eye[k, range(3), range(3), 0] = s
    W, n = links_oracle.unitarize(eye, allow_svd=False)
    assert np.abs(W - np.sign(eye) * (eye != 0)).max() < 1e-14


_LIVE_LINKS = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from milc_qcd_b200 import fields as F
from oracle.pyoracle import LinksOracle, MilcRef
dims = %(dims)r
o, r = LinksOracle(), MilcRef(dims)
for spread, tol in ((0.4, 1e-13), (5.0, 1e-9)):
    U = F.make_thin_links(dims, seed=99, spread=spread)
    a, b = o.hisq_links(dims, U), r.hisq_links(U)
    assert abs(a['nsvd'] - b['nsvd']) <= 1, (a['nsvd'], b['nsvd'])
    for k in ('V', 'W', 'fat', 'lng'):
        assert np.abs(a[k] - b[k]).max() <= tol * np.abs(b[k]).max(), (spread, k, np.abs(a[k] - b[k]).max())
    # one smearing level with other coefficients (asqtad, tadpole-improved with u0 = 0.86)
    u0 = 0.86
    c = (5.0/8.0, -1.0/(24*u0**2), -1.0/(16*u0**2), 1.0/(64*u0**4), -1.0/(384*u0**6), -1.0/(16*u0**4))
    fa, la = o.smear(dims, U, c)
    fb, lb = r.smear(U, c)
    assert np.abs(fa - fb).max() <= 1e-14 * np.abs(fb).max() and np.abs(la - lb).max() <= 1e-14 * np.abs(lb).max()
    # one link only: the reference skips the staples
    fa, _ = o.smear(dims, U, (0.125, -1/24., 0, 0, 0, 0), want_long=False)
    fb, _ = r.smear(U, (0.125, -1/24., 0, 0, 0, 0), want_long=False)
    assert np.abs(fa - fb).max() <= 1e-15
print('LIVE-OK')
"""


@pytest.mark.parametrize("dims", [(4, 6, 4, 8), (6, 4, 8, 4)])
def test_links_oracle_matches_compiled_reference_live(dims):
    from oracle.pyoracle import ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libmilcref.so not built (needs /root/reference)")
    out = subprocess.run([sys.executable, "-c", _LIVE_LINKS % dict(root=ROOT, dims=dims)], capture_output=True,
                         text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- UML propagator-solve sequence (row f3): the composition used as the GPU tests' oracle --------
def uml_from_primitives(o, dims, fat, lng, src, guess, mass, niter, nrestart, resid):
    """generic_ks/mat_invert.c:328-402 composed from the pinned oracle primitives (the same
    function body as tests/test_gpu_sequences.py::uml_oracle)."""
    h = src.shape[0] // 2
    tmp = -o.dslash(dims, fat, lng, src, EVENANDODD) + 2 * mass * src
    dst = guess.copy()
    it_e, _ = o.congrad(dims, fat, lng, tmp, dst, mass, EVEN, niter, nrestart, resid)
    ttt = o.dslash(dims, fat, lng, dst, ODD)
    dst[h:] = (src[h:] - ttt[h:]) / (2 * mass)
    it_o, _ = o.congrad(dims, fat, lng, tmp, dst, mass, ODD, niter, nrestart, resid)
    return dst, it_e + it_o


_LIVE_UML = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + '/tests')
from milc_qcd_b200 import fields as F
from oracle.pyoracle import Oracle, MilcRef, EVENANDODD
from test_oracle import uml_from_primitives
dims = (6, 6, 6, 8)
fat, lng = F.make_links(dims, seed=77)
o, r = Oracle(), MilcRef(dims)
r.set_links(fat, lng)
srcs = np.stack([F.make_source(dims, seed=200 + k, parity=EVENANDODD) for k in range(3)])
for nsrc in (1, 3):
    dsts = np.zeros_like(srcs[:nsrc])
    it, q = r.mat_invert_uml(srcs[:nsrc], dsts, 0.1, 300, 5, 1e-9)
    tot = 0
    for k in range(nsrc):
        want, itk = uml_from_primitives(o, dims, fat, lng, srcs[k], np.zeros_like(srcs[k]), 0.1, 300, 5, 1e-9)
        tot += itk
        assert np.linalg.norm(dsts[k] - want) <= 1e-8 * np.linalg.norm(want), k
    assert abs(it - tot) <= 2 * nsrc and q['converged'] == 1, (it, tot)
print('LIVE-OK')
"""


def test_uml_sequence_oracle_matches_compiled_reference_live():
    from oracle.pyoracle import ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libmilcref.so not built (needs /root/reference)")
    out = subprocess.run([sys.executable, "-c", _LIVE_UML % dict(root=ROOT)], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- low-mode deflation (row f4) -----------------------------------------------------------------
def uml_deflated_from_primitives(o, dims, fat, lng, src, mass, niter, nrestart, resid, ev, lam):
    """mat_invert_uml_field with qic->deflate = 1 (generic_ks/mat_invert.c:328-402) from oracle primitives."""
    from oracle.pyoracle import EVEN, ODD, EVENANDODD
    h = src.shape[0] // 2
    dst = np.zeros_like(src)
    tmp = -o.dslash(dims, fat, lng, src, EVENANDODD) + 2 * mass * src
    o.deflate(dims, dst, tmp, mass, ev, lam, EVEN)
    it_e, _ = o.congrad(dims, fat, lng, tmp, dst, mass, EVEN, niter, nrestart, resid)
    ttt = o.dslash(dims, fat, lng, dst, ODD)
    dst[h:] = (src[h:] - ttt[h:]) / (2 * mass)
    o.deflate(dims, dst, tmp, mass, ev, lam, ODD)
    it_o, _ = o.congrad(dims, fat, lng, tmp, dst, mass, ODD, niter, nrestart, resid)
    return dst, it_e + it_o


def _deflate_case():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_deflate import low_modes
    from milc_qcd_b200 import fields as F
    from oracle.pyoracle import Oracle, EVENANDODD
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_uml_deflated.npz"))
    dims = tuple(int(d) for d in g["dims"])
    fat, lng = F.make_links(dims, seed=int(g["link_seed"]))
    o = Oracle()
    lam, ev = low_modes(o, dims, fat, lng)
    src = F.make_source(dims, seed=int(g["src_seed"]), parity=EVENANDODD)
    return g, dims, fat, lng, o, lam, ev, src


def test_deflation_oracle_matches_reference_golden():
    """kso_deflate inside the UML sequence against the committed output of the reference's
    mat_invert_uml_field with qic->deflate = 1 (tests/golden/make_golden_deflate.py): same iteration counts
    (the CG trajectory depends on the trial solution) and solutions, for 8, 48 and all 384 low modes."""
    g, dims, fat, lng, o, lam, ev, src = _deflate_case()
    assert np.abs(lam[:48] - g["eigval"]).max() <= 1e-12 * lam[47]
    args = (float(g["mass"]), int(g["niter"]), int(g["nrestart"]), float(g["resid"]))
    for k, it_ref, want in zip(g["nvecs"], g["iters"], g["solutions"]):
        got, it = uml_deflated_from_primitives(o, dims, fat, lng, src, *args, ev[:k], lam[:k])
        assert abs(it - int(it_ref)) <= 2, (k, it, it_ref)
        assert np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want), k
    assert int(g["iters"][-1]) == 2 and int(g["iters"][0]) < int(g["iters_plain"]) // 2
    # orthonormal input: projecting out and adding back are independent of the order of the vectors
    from oracle.pyoracle import EVEN
    rng = np.random.default_rng(4)
    d1 = rng.standard_normal(src.shape)
    d2 = d1.copy()
    perm = rng.permutation(16)
    o.deflate(dims, d1, src, 0.1, ev[:16], lam[:16], EVEN)
    o.deflate(dims, d2, src, 0.1, ev[:16][perm], lam[:16][perm], EVEN)
    assert np.abs(d1 - d2).max() <= 1e-12 * np.abs(d1).max()


def test_deflation_reference_live_matches_golden():
    from oracle.pyoracle import ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libmilcref.so not built (needs /root/reference)")
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from test_oracle import _deflate_case
from oracle.pyoracle import MilcRef
g, dims, fat, lng, o, lam, ev, src = _deflate_case()
r = MilcRef(dims)
r.set_links(fat, lng)
for k, it_ref, want in zip(g['nvecs'], g['iters'], g['solutions']):
    dst = np.zeros_like(src)
    it, q = r.mat_invert_uml_deflated(src, dst, float(g['mass']), int(g['niter']), int(g['nrestart']), float(g['resid']), ev[:k], lam[:k])
    assert abs(it - int(it_ref)) <= 1 and q['converged'] == 1, (k, it, it_ref)
    assert np.linalg.norm(dst - want) <= 1e-10 * np.linalg.norm(want), k
print('LIVE-OK')
""" % (ROOT, os.path.dirname(__file__))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- HISQ fermion force (row f2): the oracle for the next row, pinned ahead of the CUDA work ------
def _generator(a):
    T = np.zeros((3, 3), complex)
    if a == 0:
        T[0, 1] = T[1, 0] = 1
    elif a == 1:
        T[0, 1], T[1, 0] = -1j, 1j
    elif a == 2:
        T[0, 0], T[1, 1] = 1, -1
    elif a == 3:
        T[0, 2] = T[2, 0] = 1
    elif a == 4:
        T[1, 2], T[2, 1] = -1j, 1j
    else:
        T[0, 0] = T[1, 1] = 1 / np.sqrt(3)
        T[2, 2] = -2 / np.sqrt(3)
    return T


def _ahmat(m):
    """anti_hermitmat -> 3x3 anti-Hermitian matrix (libraries/uncmp_ahmat.c)."""
    A = np.zeros((3, 3), complex)
    A[0, 0], A[1, 1], A[2, 2] = 1j * m[6], 1j * m[7], 1j * m[8]
    A[0, 1], A[1, 0] = m[0] + 1j * m[1], -m[0] + 1j * m[1]
    A[0, 2], A[2, 0] = m[2] + 1j * m[3], -m[2] + 1j * m[3]
    A[1, 2], A[2, 1] = m[4] + 1j * m[5], -m[4] + 1j * m[5]
    return A


def test_reference_force_golden_is_the_derivative_of_the_oracle_action(oracle, links_oracle):
    """The committed output of the reference's eo_fermion_force_multi (tests/golden/
    make_golden_force.py) against an independent finite-difference derivative built from the
    pinned oracles: with S(U) = sum_j res_j |D_oe[U] X_j|^2 (HISQ links from ks_links_oracle.c,
    stencil from ks_oracle.c) and U_mu(x) -> exp(i t T) U_mu(x), the reference's momentum update A
    (eps = 1) satisfies dS/dt = -Re tr(i T A).  Fixes the convention the CUDA force has to meet."""
    from scipy.linalg import expm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_force.npz"))
    dims = tuple(int(d) for d in g["dims"])
    U, X, res, mom = g["U"], g["multi_x"], g["residues"], g["mom"]
    h = U.shape[0] // 2

    def action(Ur):
        L = links_oracle.hisq_links(dims, Ur)
        return sum(res[j] * np.sum(oracle.dslash(dims, L["fat"], L["lng"], X[j], ODD)[h:] ** 2) for j in range(len(res)))

    Uc = U[..., 0] + 1j * U[..., 1]
    t = 1e-5
    for (i, mu, a) in [(17, 2, 2), (40, 3, 3), (100, 0, 4), (201, 1, 0)]:
        T = _generator(a)
        s = []
        for sgn in (1, -1):
            Up = Uc.copy()
            Up[i, mu] = expm(1j * sgn * t * T) @ Uc[i, mu]
            s.append(action(np.ascontiguousarray(np.stack([Up.real, Up.imag], axis=-1))))
        fd = (s[0] - s[1]) / (2 * t)
        want = -np.trace(1j * T @ _ahmat(mom[i, mu])).real
        assert abs(fd - want) <= 1e-6 * max(1.0, abs(want)), (i, mu, a, fd, want)
    # traceless and anti-Hermitian by construction of the packed format; zero "space" member
    assert np.abs(mom[..., 6] + mom[..., 7] + mom[..., 8]).max() < 1e-12


@pytest.mark.parametrize("name,tol", [("ref_hisq_force.npz", 1e-11), ("ref_hisq_force_rough.npz", 1e-9)])
def test_force_oracle_matches_reference_golden(links_oracle, name, tol):
    """ks_force_oracle.c against the committed output of the reference's eo_fermion_force_multi.  The rough
    file has 11 links on the reference's HISQ_FORCE_FILTER / SVD branches: without the filter the
    restatement is off by O(1) there.  (1e-9: the reference's eigenvalues come from the closed-form cubic.)"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    dims = tuple(int(d) for d in g["dims"])
    scale = np.abs(g["mom"]).max()
    mom = links_oracle.hisq_force(dims, g["U"], g["multi_x"], g["residues"], float(g["eps"]))
    assert np.abs(mom - g["mom"]).max() <= tol * scale
    raw = links_oracle.hisq_force(dims, g["U"], g["multi_x"], g["residues"], float(g["eps"]), force_filter=0.0)
    if int(g["nsvd"]) == 0:
        assert np.array_equal(raw, mom)
    else:
        assert np.abs(raw - g["mom"]).max() > 0.1 * scale


def test_force_oracle_with_naik_epsilons_matches_reference_golden(links_oracle):
    """Several Naik epsilons (fermion_force_hisq_multi.c:1285-1375): five terms in three classes,
    tests/golden/ref_hisq_force_naik.npz from the reference's eo_fermion_force_multi with n_naiks = 3."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_force_naik.npz"))
    dims = tuple(int(d) for d in g["dims"])
    scale = np.abs(g["mom"]).max()
    n_orders, eps_naik = [int(v) for v in g["n_orders"]], [float(v) for v in g["eps_naik"]]
    mom = links_oracle.hisq_force_naik(dims, g["U"], g["multi_x"], g["residues"], n_orders, eps_naik, float(g["eps"]))
    assert np.abs(mom - g["mom"]).max() <= 1e-11 * scale
    plain = links_oracle.hisq_force(dims, g["U"], g["multi_x"], g["residues"], float(g["eps"]))
    assert np.abs(plain - g["mom"]).max() > 1e-3 * scale
    # one class is the plain force
    one = links_oracle.hisq_force_naik(dims, g["U"], g["multi_x"], g["residues"], [5], [0.0], float(g["eps"]))
    assert np.array_equal(one, plain)


def test_force_filter_is_the_derivative_of_the_shifted_projection(links_oracle):
    """What HISQ_FORCE_FILTER does to one link (generic_ks/su3_mat_op.c:1680-1734): when the smallest eigenvalue
    of Q = V^+ V is below the filter, the reverse step is the exact derivative of V (Q + filter)^-1/2, by
    finite differences; above it, of V Q^-1/2."""
    rng = np.random.default_rng(12)

    def w(Vc, shift):
        gq, E = np.linalg.eigh(Vc.conj().T @ Vc + shift * np.eye(3))
        return Vc @ (E * gq ** -0.5) @ E.conj().T

    for smin, shifted in [(3e-3, True), (2e-2, False)]:          # g_min = 9e-6 and 4e-4
        A = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
        u, _, vh = np.linalg.svd(A)
        Vc = u @ np.diag([1.1, 0.8, smin]) @ vh
        GW = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
        dV = (rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))) * 1e-4 * smin
        GV = links_oracle.unitarize_bwd(np.stack([Vc.real, Vc.imag], -1), np.stack([GW.real, GW.imag], -1), 5e-5)
        GV = GV[..., 0] + 1j * GV[..., 1]
        pred = np.trace(GV.conj().T @ dV).real
        fd = {}
        for shift in (0.0, 5e-5):
            fd[shift] = 0.5 * np.trace(GW.conj().T @ (w(Vc + dV, shift) - w(Vc - dV, shift))).real
        assert abs(pred - fd[5e-5 if shifted else 0.0]) <= 1e-5 * abs(pred)
        assert abs(pred - fd[0.0 if shifted else 5e-5]) > 1e-3 * abs(pred)


def test_reference_force_live_matches_golden():
    from oracle.pyoracle import ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libmilcref.so not built (needs /root/reference)")
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from oracle.pyoracle import MilcRef
g = np.load(%r)
ref = MilcRef(tuple(int(d) for d in g['dims']))
mom, n = ref.hisq_force(g['U'], g['multi_x'], g['residues'], float(g['eps']))
assert np.abs(mom - g['mom']).max() <= 1e-12 * np.abs(g['mom']).max()
m2, _ = ref.hisq_force(g['U'], g['multi_x'], 2.0 * g['residues'], 0.5)      # linear in eps * residues
assert np.abs(m2 - g['mom']).max() <= 1e-12 * np.abs(g['mom']).max()
g = np.load(%r)                                                              # filter / SVD branches
mom, n = ref.hisq_force(g['U'], g['multi_x'], g['residues'], float(g['eps']))
assert n == int(g['nsvd']) and n > 0
assert np.abs(mom - g['mom']).max() <= 1e-12 * np.abs(g['mom']).max()
g = np.load(%r)                                                              # three Naik-epsilon classes
mom, n = ref.hisq_force_naik(g['U'], g['multi_x'], g['residues'], [int(v) for v in g['n_orders']],
                             [float(v) for v in g['eps_naik']], float(g['eps']))
assert np.abs(mom - g['mom']).max() <= 1e-12 * np.abs(g['mom']).max()
print('LIVE-OK')
""" % (ROOT, os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_force.npz"),
       os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_force_rough.npz"),
       os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_force_naik.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "LIVE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

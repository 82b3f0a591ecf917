"""GPU parity tests of the fermion-link construction: the CUDA
smearing / U(3) projection / Naik kernels through the C ABI against the CPU oracle
(oracle/ks_links_oracle.c, pinned on the reference's compiled chain) and against the committed
output of the reference's own create_hisq_links_milc (tests/golden/ref_hisq_links.npz)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_LINKS = os.path.join(os.path.dirname(__file__), "golden", "ref_hisq_links.npz")


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def api():
    from milc_qcd_b200 import api
    yield api
    api.finalize()


@pytest.fixture(scope="module")
def links_oracle():
    from oracle.pyoracle import LinksOracle
    return LinksOracle()


@pytest.mark.parametrize("tag,tol", [("smooth", 1e-13), ("rough", 1e-9)])
def test_hisq_chain_matches_reference_golden(api, tag, tol):
    from milc_qcd_b200 import fields as F
    g = np.load(GOLDEN_LINKS)
    dims = tuple(int(d) for d in g["dims"])
    U = F.make_thin_links(dims, seed=4321, spread=float(g[tag + "_spread"]))
    ctx = api.Context(dims)
    out = ctx.hisq_links(U, g["coeffs"][0], g["coeffs"][1])
    assert abs(out["nsvd"] - int(g[tag + "_nsvd"])) <= 1
    assert rel_err(out["V"], g[tag + "_V"]) <= 1e-14
    for k in ("W", "fat", "lng"):
        assert rel_err(out[k], g[tag + "_" + k]) <= tol, k
    # the two entry points MILC's glue binds (qudaLoadUnitarizedLink, qudaLoadKSLink)
    V, W, n = ctx.unitarized_links(U, g["coeffs"][0])
    assert np.array_equal(V, out["V"]) and np.array_equal(W, out["W"]) and n == out["nsvd"]
    fat, lng = ctx.ks_links(W, g["coeffs"][1])
    assert np.array_equal(fat, out["fat"]) and np.array_equal(lng, out["lng"])
    ctx.close()


@pytest.mark.parametrize("dims", [(4, 6, 4, 8), (8, 8, 8, 8), (6, 4, 2, 10), (2, 2, 2, 2)])
@pytest.mark.parametrize("spread", [0.4, 5.0])
def test_smearing_and_projection_match_oracle(api, links_oracle, dims, spread):
    from milc_qcd_b200 import fields as F
    U = F.make_thin_links(dims, seed=17, spread=spread)
    ctx = api.Context(dims)
    u0 = 0.86   # tadpole-improved asqtad coefficients (generic_ks/imp_actions/asqtad_action.h)
    asqtad = (5.0 / 8.0, -1.0 / (24 * u0 ** 2), -1.0 / (16 * u0 ** 2), 1.0 / (64 * u0 ** 4), -1.0 / (384 * u0 ** 6),
              -1.0 / (16 * u0 ** 4))
    for coeffs in (ctx.HISQ_FAT7, ctx.HISQ_ASQTAD_LIKE, asqtad, (0.125, -1.0 / 24.0, 0, 0, 0, 0)):
        fat, lng = ctx.ks_links(U, coeffs)
        fo, lo = links_oracle.smear(dims, U, coeffs)
        assert rel_err(fat, fo) <= 1e-13 and rel_err(lng, lo) <= 1e-13, coeffs
    fat, none = ctx.ks_links(U, ctx.HISQ_FAT7, want_long=False)
    assert none is None and rel_err(fat, links_oracle.smear(dims, U, ctx.HISQ_FAT7)[0]) <= 1e-13
    o = links_oracle.hisq_links(dims, U)
    out = ctx.hisq_links(U)
    assert abs(out["nsvd"] - o["nsvd"]) <= max(1, o["nsvd"] // 10)
    tol = 1e-13 if spread < 1 else 1e-8
    for k in ("V", "W", "fat", "lng"):
        assert rel_err(out[k], o[k]) <= (1e-13 if k == "V" else tol), (k, rel_err(out[k], o[k]))
    W = out["W"][..., 0] + 1j * out["W"][..., 1]
    assert np.abs(W @ np.conj(np.swapaxes(W, -1, -2)) - np.eye(3)).max() <= 1e-9
    ctx.close()


def test_float_hosts_and_solver_handoff(api, links_oracle, oracle):
    """MILC_PRECISION=1 callers hand over float links (the device still works in double); links
    built on the GPU feed the stencil like host-built ones."""
    from milc_qcd_b200 import fields as F
    dims = (4, 6, 4, 8)
    U = F.make_thin_links(dims, seed=23, spread=0.4)
    ctx = api.Context(dims)
    o = links_oracle.hisq_links(dims, U)
    out = ctx.hisq_links(U.astype(np.float32))
    for k in ("W", "fat", "lng"):
        assert out[k].dtype == np.float32 and rel_err(out[k].astype(np.float64), o[k]) <= 5e-6
    out = ctx.hisq_links(U)
    ctx.load_links(out["fat"], out["lng"])
    assert ctx.long_link_info()[0] == 7          # c_naik * W W W with W in U(3): two rows + factor suffice
    src = F.make_source(dims, seed=3, parity=3)
    got = np.zeros_like(src)
    ctx.dslash(src, got, 3)
    want = oracle.dslash(dims, o["fat"], o["lng"], src, 3)
    assert rel_err(got, want) <= 1e-12
    ctx.close()


def test_device_generated_chain_is_reproducible(api, links_oracle):
    """The benchmark face: Haar-random thin links made on the device, chain timed with everything
    resident; its fields read back and checked against the oracle on the same input."""
    dims = (8, 8, 8, 8)
    ctx = api.Context(dims)
    ms, nsvd = ctx.hisq_links_time(1234, 2)
    assert ms > 0
    U = ctx.hisq_links_fetch(0)
    Uc = U[..., 0] + 1j * U[..., 1]
    assert np.abs(Uc @ np.conj(np.swapaxes(Uc, -1, -2)) - np.eye(3)).max() <= 1e-12   # phases are signs
    o = links_oracle.hisq_links(dims, U)
    assert abs(nsvd - o["nsvd"]) <= max(1, o["nsvd"] // 10)
    for which, k in ((1, "V"), (2, "W"), (3, "fat"), (4, "lng")):
        got = ctx.hisq_links_fetch(which)
        assert rel_err(got, o[k]) <= (1e-13 if k == "V" else 1e-7), k
    ctx.close()

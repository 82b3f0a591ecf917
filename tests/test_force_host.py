"""CPU test of the CUDA force's arithmetic: the __host__ __device__
site routines of milc_qcd_b200/csrc/force.cuh and the chain that sequences them, run in host loops
(tests/host/force_host.cu, compiled with nvcc for the host), against the CPU oracle
(oracle/ks_force_oracle.c) and the committed output of the reference's eo_fermion_force_multi."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libforce_host.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host_force():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    src = os.path.join(HOST_DIR, "force_host.cu")
    hdr = os.path.join(ROOT, "milc_qcd_b200", "csrc", "force.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-x", "cu", "--shared", "-Xcompiler", "-fPIC", "-o", SO, src])
    lib = C.CDLL(SO)
    lib.force_host.restype = None
    lib.force_host.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, _dp]
    return lib


def _run(lib, dims, U, V, W, X, c1, c3, eps, naik_in_oprod, coeffs1, coeffs2, force_filter=5.0e-5, n_naik_terms=0, split=False):
    mom = np.zeros((U.shape[0], 4, 10))
    lib.force_host(np.ascontiguousarray(dims, np.int32), np.ascontiguousarray(coeffs1, np.float64),
                   np.ascontiguousarray(coeffs2, np.float64), np.ascontiguousarray(U), np.ascontiguousarray(V),
                   np.ascontiguousarray(W), np.ascontiguousarray(X), np.ascontiguousarray(c1, np.float64),
                   np.ascontiguousarray(c3, np.float64), X.shape[0], n_naik_terms, eps, int(naik_in_oprod), force_filter, int(split), mom)
    return mom


def test_force_site_routines_match_reference_golden(host_force):
    from oracle.pyoracle import LinksOracle
    lo = LinksOracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_hisq_force.npz"))
    dims = tuple(int(d) for d in g["dims"])
    U, X, res = g["U"], g["multi_x"], g["residues"]
    L = lo.hisq_links(dims, U)
    naik = lo.ASQTAD_LIKE[1]
    # the seam's convention (qudaHisqForce): one-hop 2 res, three-hop naik * 2 res
    mom = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE)
    assert np.abs(mom - g["mom"]).max() <= 1e-11 * np.abs(g["mom"]).max()
    # the other convention: Naik coefficient applied in the chain
    mom2 = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, 2 * res, float(g["eps"]), False, lo.FAT7, lo.ASQTAD_LIKE)
    assert np.abs(mom2 - g["mom"]).max() <= 1e-11 * np.abs(g["mom"]).max()


@pytest.mark.parametrize("dims,spread", [((4, 6, 2, 4), 0.4), ((2, 2, 4, 6), 0.8), ((8, 4, 4, 6), 0.6)])
def test_force_site_routines_match_oracle(host_force, dims, spread):
    """Asymmetric lattices (incl. extents of 2, where +mu and -mu are the same neighbour), other
    coefficients (tadpole-improved asqtad levels), three terms."""
    from milc_qcd_b200 import fields as F
    from oracle.pyoracle import LinksOracle, Oracle, ODD
    lo, o = LinksOracle(), Oracle()
    V_ = int(np.prod(dims))
    h = V_ // 2
    U = F.make_thin_links(dims, seed=5, spread=spread)
    u0 = 0.9
    c2 = (1.0, -1.0 / (24 * u0 ** 2), -1.0 / (16 * u0 ** 2), 1.0 / (64 * u0 ** 4), -1.0 / (384 * u0 ** 6), -1.0 / (8 * u0 ** 4))
    L = lo.hisq_links(dims, U, lo.FAT7, c2, allow_svd=False)
    rng = np.random.default_rng(8)
    X = rng.standard_normal((3, V_, 3, 2))
    X[:, h:] = 0
    for j in range(3):
        X[j, h:] = o.dslash(dims, L["fat"], L["lng"], X[j], ODD)[h:]
    res = np.array([0.4, -1.1, 0.05])
    want = lo.hisq_force(dims, U, X, res, 0.3, lo.FAT7, c2)
    got = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, c2[1] * 2 * res, 0.3, True, lo.FAT7, c2)
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()


def test_force_filter_matches_reference_on_rough_links(host_force):
    """Links rough enough that eigenvalues of V^+ V fall below the reference's HISQ_FORCE_FILTER (5e-5)
    and its SVD thresholds: the committed output of the reference's eo_fermion_force_multi
    (tests/golden/ref_hisq_force_rough.npz, 11 links on the filter / SVD branches).  The reference's
    own eigenvalues come from the closed-form cubic, hence 1e-9 rather than 1e-11."""
    from oracle.pyoracle import LinksOracle
    lo = LinksOracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_hisq_force_rough.npz"))
    dims = tuple(int(d) for d in g["dims"])
    U, X, res = g["U"], g["multi_x"], g["residues"]
    assert int(g["nsvd"]) > 0
    L = lo.hisq_links(dims, U)
    naik = lo.ASQTAD_LIKE[1]
    scale = np.abs(g["mom"]).max()
    mom = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE, 5.0e-5)
    assert np.abs(mom - g["mom"]).max() <= 1e-9 * scale
    assert np.abs(mom - lo.hisq_force(dims, U, X, res, float(g["eps"]))).max() <= 1e-11 * scale
    # the split form of the backward staple passes (ForceBufs::split)
    two = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE,
               5.0e-5, split=True)
    assert np.abs(two - mom).max() <= 1e-12 * scale
    # the two-role form of the full passes (ForceBufs::pair): same products, same order of summation
    pair = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE,
                5.0e-5, split=2)
    assert np.array_equal(pair, mom)
    # the filter is what makes the difference on this input, and switching it off matches the oracle's unfiltered force
    raw = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE, 0.0)
    assert np.abs(raw - g["mom"]).max() > 0.1 * scale
    want = lo.hisq_force(dims, U, X, res, float(g["eps"]), force_filter=0.0)
    assert np.abs(raw - want).max() <= 1e-10 * np.abs(want).max()


def _naik_weights(res, n_orders, eps_naik, c3):
    """The seam's coeff[num_terms + i] (fermion_force_hisq_multi.c:2205-2213): one- and three-hop weights of
    the terms solved with a Naik epsilon."""
    one, three, j = [], [], n_orders[0]
    for k in range(1, len(n_orders)):
        for _ in range(n_orders[k]):
            one.append(c3[0] * eps_naik[k] * 2 * res[j])
            three.append(c3[1] * eps_naik[k] * 2 * res[j])
            j += 1
    return np.array(one), np.array(three)


def test_force_with_naik_epsilons_matches_reference_golden(host_force):
    """Five terms in three Naik-epsilon classes (tests/golden/ref_hisq_force_naik.npz, from the reference's
    eo_fermion_force_multi with n_naiks = 3)."""
    from oracle.pyoracle import LinksOracle
    lo = LinksOracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_hisq_force_naik.npz"))
    dims = tuple(int(d) for d in g["dims"])
    U, X, res = g["U"], g["multi_x"], g["residues"]
    n_orders, eps_naik = [int(v) for v in g["n_orders"]], [float(v) for v in g["eps_naik"]]
    L = lo.hisq_links(dims, U)
    naik = lo.ASQTAD_LIKE[1]
    one, three = _naik_weights(res, n_orders, eps_naik, lo.NAIK_TABLE)
    c1 = np.concatenate([2 * res, one])
    c3 = np.concatenate([naik * 2 * res, three])
    mom = _run(host_force, dims, U, L["V"], L["W"], X, c1, c3, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE,
               n_naik_terms=len(one))
    scale = np.abs(g["mom"]).max()
    assert np.abs(mom - g["mom"]).max() <= 1e-11 * scale
    assert np.abs(mom - lo.hisq_force_naik(dims, U, X, res, n_orders, eps_naik, float(g["eps"]))).max() <= 1e-11 * scale
    # the epsilons matter on this input
    plain = _run(host_force, dims, U, L["V"], L["W"], X, 2 * res, naik * 2 * res, float(g["eps"]), True, lo.FAT7, lo.ASQTAD_LIKE)
    assert np.abs(plain - g["mom"]).max() > 1e-3 * scale


@pytest.mark.parametrize("Vh", [1, 24, 63, 64, 65, 96, 648, 1 << 14])
def test_interleaved_launch_order_visits_every_site_once(host_force, Vh):
    """common.cuh interleaved_site (launch order of the staple passes of the link construction and of the force):
    a bijection from the live launch indices onto the 2 Vh sites, parity uniform within every warp, and the even and
    the odd sites of one CTA cover the same checkerboard range."""
    cap = (Vh // 64 + 2) * 128
    out = np.full(cap, -7, dtype=np.int32)
    host_force.interleaved_order.restype = C.c_int
    host_force.interleaved_order.argtypes = [C.c_int, _ip, C.c_int]
    n = host_force.interleaved_order(Vh, out, cap)
    assert n % 128 == 0 and n <= cap and n >= 2 * Vh
    got = out[:n]
    live = got[got >= 0]
    assert sorted(live.tolist()) == list(range(2 * Vh))
    for w in range(n // 32):
        lane = got[32 * w:32 * w + 32]
        lane = lane[lane >= 0]
        assert len(set((lane >= Vh).tolist())) <= 1
    for b in range(n // 128):
        cta = got[128 * b:128 * b + 128]
        ev, od = cta[:64], cta[64:]
        assert np.all((ev < Vh)) and np.all((od < 0) | (od >= Vh))
        assert np.array_equal(ev[ev >= 0], od[od >= 0] - Vh)

"""CPU test of the route-1 shim's host logic (milc_qcd_b200/csrc_milc/milc_shim.c): the shim compiled against
a recording stand-in for the b200ks C ABI (tests/host/b200ks_stub.c).  Checks what the shim does around the
solver calls, as the reference's glue does (generic_ks/d_congrad5_fn_gpu.c:35-172, ks_multicg_offset_gpu.c:38-252):
qic bookkeeping, the zero-source shortcut, when the links are (re-)uploaded, the eigenvector hand-over and the
qic->deflate switch.  No physics: the stub's "solvers" copy the source."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libmilc_shim_stub.so")
EVEN, ODD, EVENANDODD = 2, 1, 3
DIMS = (4, 4, 2, 2)
V = int(np.prod(DIMS))


class Qic(C.Structure):          # include/b200ks_milc.h (= include/generic_quark_types.h:167-190), PRECISION 2
    _fields_ = [("prec", C.c_int), ("min", C.c_int), ("max", C.c_int), ("nrestart", C.c_int), ("parity", C.c_int),
                ("start_flag", C.c_int), ("nsrc", C.c_int), ("deflate", C.c_int), ("resid", C.c_double),
                ("relresid", C.c_double), ("mixed_rsq", C.c_double), ("final_rsq", C.c_double), ("final_relrsq", C.c_double),
                ("size_r", C.c_double), ("size_relr", C.c_double), ("converged", C.c_int), ("final_iters", C.c_int),
                ("final_restart", C.c_int), ("inv_type", C.c_int), ("mgparamfile", C.c_char * 256)]


class KsParam(C.Structure):      # include/generic_quark_types.h:131-139
    _fields_ = [("mass", C.c_double), ("charge", C.c_double), ("offset", C.c_double), ("residue", C.c_double),
                ("naik_term_epsilon_index", C.c_int), ("charge_index", C.c_int), ("naik_term_epsilon", C.c_double)]


class FnLinks(C.Structure):      # include/fn_links.h:12-20
    _fields_ = [("phase", C.c_void_p), ("fat", C.c_void_p), ("lng", C.c_void_p), ("fatback", C.c_void_p),
                ("lngback", C.c_void_p), ("eps_naik", C.c_double), ("notify_quda_new_links", C.c_int)]


@pytest.fixture()
def shim():
    srcs = [os.path.join(ROOT, "milc_qcd_b200", "csrc_milc", "milc_shim.c"), os.path.join(HOST_DIR, "b200ks_stub.c")]
    deps = srcs + [os.path.join(ROOT, "include", f) for f in ("b200ks.h", "b200ks_milc.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in deps):
        # -Bsymbolic: the shim's calls bind to the stub in this .so even if the real libb200ks.so is already
        # loaded globally in the test process (it would fail without a GPU and terminate the run, as designed)
        subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", "-std=gnu99", "-Wall", "-Wl,-Bsymbolic", "-DMILC_PRECISION=2",
                               "-I", os.path.join(ROOT, "include"), "-o", SO] + srcs)
    lib = C.CDLL(SO)
    lib.stub_log.restype = C.c_char_p
    lib.b200ks_milc_setup(*DIMS, 0)
    lib.stub_reset()
    yield lib
    lib.b200ks_milc_finalize()


def _log(lib):
    out = lib.stub_log().decode()
    lib.stub_reset()
    return [ln for ln in out.splitlines() if ln]


def _fn(fat, lng):
    return FnLinks(None, fat.ctypes.data, lng.ctypes.data, None, None, 0.0, 1)


def _qic(parity, **kw):
    q = Qic(prec=2, min=0, max=300, nrestart=5, parity=parity, start_flag=0, nsrc=1, deflate=0, resid=1e-9, relresid=0.0)
    for k, v in kw.items():
        setattr(q, k, v)
    return q


def test_single_mass_solve_bookkeeping_and_link_cache(shim):
    rng = np.random.default_rng(1)
    fat, lng = rng.standard_normal((V, 4, 3, 3, 2)), rng.standard_normal((V, 4, 3, 3, 2))
    fn = _fn(fat, lng)
    src, dst = rng.standard_normal((V, 3, 2)), np.zeros((V, 3, 2))
    src[V // 2:] = 0
    q = _qic(EVEN)
    shim.ks_congrad_parity_gpu.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Qic), C.c_double, C.POINTER(FnLinks)]
    it = shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.05, C.byref(fn))
    log = _log(shim)
    assert it == 17 and q.final_iters == 17 and q.converged == 1 and q.final_restart == 1
    assert q.final_rsq == 1e-20 and q.size_r == 2e-20
    assert log[0].startswith("create 4 4 2 2") and log[1].startswith("load_links prec 2")
    assert "congrad mass 0.05 parity 2 max 300 nrestart 5 resid 1e-09" in log[2]
    assert fn.notify_quda_new_links == 0 and shim.b200ks_milc_total_iters() == 17
    assert np.array_equal(dst[:V // 2], src[:V // 2])
    # same links again: no upload
    shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.05, C.byref(fn))
    assert not any(ln.startswith("load_links") for ln in _log(shim))
    # edited in place without notice (boundary_twist_fn): the fingerprint sees it
    fat[3, 1, 0, 0, 0] += 1.0
    shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.05, C.byref(fn))
    assert any(ln.startswith("load_links") for ln in _log(shim))
    # MILC's own notification
    fn.notify_quda_new_links = 1
    shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.05, C.byref(fn))
    assert any(ln.startswith("load_links") for ln in _log(shim)) and fn.notify_quda_new_links == 0
    # zero source on the solve's parity: zero solution, no solver call (d_congrad5_fn_gpu.c:63-89)
    q2 = _qic(ODD)
    dst[:] = 7.0
    it = shim.ks_congrad_parity_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q2), 0.05, C.byref(fn))
    assert it == 0 and not _log(shim)
    assert np.all(dst[V // 2:] == 0) and np.all(dst[:V // 2] == 7.0)


def test_multishift_bookkeeping(shim):
    rng = np.random.default_rng(2)
    fat, lng = rng.standard_normal((V, 4, 3, 3, 2)), rng.standard_normal((V, 4, 3, 3, 2))
    fn = _fn(fat, lng)
    src = rng.standard_normal((V, 3, 2))
    n = 3
    psim = [np.ones((V, 3, 2)) for _ in range(n)]
    pp = (C.c_void_p * n)(*[p.ctypes.data for p in psim])
    ksp = (KsParam * n)()
    for j, off in enumerate((0.01, 0.04, 0.25)):
        ksp[j].offset = off
    qic = (Qic * n)(*[_qic(ODD) for _ in range(n)])
    shim.ks_multicg_offset_field_gpu.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(KsParam), C.c_int, C.POINTER(Qic),
                                                 C.POINTER(FnLinks)]
    it = shim.ks_multicg_offset_field_gpu(src.ctypes.data, pp, ksp, n, qic, C.byref(fn))
    log = _log(shim)
    assert it == 23 and any("multicg n 3 parity 1 offsets 0.01 0.04 0.25" in ln for ln in log)
    assert all(qic[j].final_iters == 23 and qic[j].converged == 1 for j in range(n))
    assert all(np.array_equal(p[V // 2:], src[V // 2:]) and np.all(p[:V // 2] == 1.0) for p in psim)
    assert shim.ks_multicg_offset_field_gpu(src.ctypes.data, pp, ksp, 0, qic, C.byref(fn)) == 0


def test_eigenvector_hand_over_and_deflate_switch(shim):
    """b200ks_milc_set_eigenvectors uploads every vector once (both parities) and declares the set; the UML
    sequences switch the deflation on exactly when qic->deflate is set and a set exists
    (generic_ks/mat_invert.c:341-353,428-437); replacing the set frees the old vectors."""
    rng = np.random.default_rng(3)
    fat, lng = rng.standard_normal((V, 4, 3, 3, 2)), rng.standard_normal((V, 4, 3, 3, 2))
    fn = _fn(fat, lng)
    src, dst = rng.standard_normal((V, 3, 2)), np.zeros((V, 3, 2))
    shim.mat_invert_uml_field_gpu.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Qic), C.c_double, C.POINTER(FnLinks)]
    q = _qic(EVENANDODD, deflate=1)
    it = shim.mat_invert_uml_field_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.1, C.byref(fn))
    log = _log(shim)
    assert it == 33 and q.final_iters == 33 and q.parity == ODD
    assert "eig_use_in_uml 0" in log and np.array_equal(dst, src)       # deflate asked for, but no vectors yet
    nv = 3
    ev = [rng.standard_normal((V, 3, 2)) for _ in range(nv)]
    evp = (C.c_void_p * nv)(*[e.ctypes.data for e in ev])
    lam = (C.c_double * nv)(1e-4, 2e-4, 5e-4)
    shim.b200ks_milc_set_eigenvectors(nv, evp, lam)
    log = _log(shim)
    assert log[0] == "eig_set n 0 uml 0"
    assert [ln for ln in log if ln.startswith("vec_upload")] == \
        ["vec_upload %d parity 3 prec 2 first %g" % (j, ev[j].ravel()[0]) for j in range(nv)]
    assert log[-1] == "eig_set n 3 uml 0 (0 0.0001) (1 0.0002) (2 0.0005)" and shim.stub_live_vecs() == nv
    shim.mat_invert_uml_field_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.1, C.byref(fn))
    assert "eig_use_in_uml 1" in _log(shim)
    q.deflate = 0
    shim.mat_invert_uml_field_gpu(src.ctypes.data, dst.ctypes.data, C.byref(q), 0.1, C.byref(fn))
    assert "eig_use_in_uml 0" in _log(shim)
    # a new set replaces the old one: old vectors freed after the set was dropped
    shim.b200ks_milc_set_eigenvectors(1, evp, lam)
    log = _log(shim)
    assert log[0] == "eig_set n 0 uml 0" and log[1:4] == ["vec_free 0", "vec_free 1", "vec_free 2"]
    assert shim.stub_live_vecs() == 1
    shim.b200ks_milc_set_eigenvectors(0, None, None)
    assert shim.stub_live_vecs() == 0


def test_incremental_eigcg_bookkeeping(shim):
    """ks_inc_eigCG_parity_gpu / calc_eigenpairs_gpu keep MILC's eigcg_params, eigVec[] and H in step with the device
    (generic_ks/inc_eigcg.c:851-950, 282-300): new sequence when Nvecs_curr == 0, only the NEW vectors are copied
    back, Nvecs shrinks when the set is nearly full, H arrives in MILC's column-major layout."""
    class DC(C.Structure):
        _fields_ = [("real", C.c_double), ("imag", C.c_double)]

    class EigcgParams(C.Structure):
        _fields_ = [("m", C.c_int), ("Nvecs", C.c_int), ("Nvecs_curr", C.c_int), ("Nvecs_max", C.c_int), ("H", C.POINTER(DC))]

    rng = np.random.default_rng(2)
    fat, lng = rng.standard_normal((V, 4, 3, 3, 2)), rng.standard_normal((V, 4, 3, 3, 2))
    fn = _fn(fat, lng)
    src, dst = rng.standard_normal((V, 3, 2)), np.zeros((V, 3, 2))
    nmax = 10
    vecs = [np.zeros((V, 3, 2)) for _ in range(nmax + 20)]
    pv = (C.c_void_p * len(vecs))(*[v.ctypes.data for v in vecs])
    val = np.zeros(len(vecs))
    ep = EigcgParams(20, 4, 0, nmax, None)
    shim.ks_inc_eigCG_parity_gpu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(EigcgParams),
                                             C.POINTER(Qic), C.c_double, C.POINTER(FnLinks)]
    shim.calc_eigenpairs_gpu.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(EigcgParams), C.c_int]
    expect = [(4, 4), (8, 2), (10, 0)]
    for call, (ncurr, nvecs_next) in enumerate(expect):
        q = _qic(EVEN)
        it = shim.ks_inc_eigCG_parity_gpu(src.ctypes.data, dst.ctypes.data, val.ctypes.data, pv, C.byref(ep), C.byref(q), 0.05,
                                          C.byref(fn))
        log = _log(shim)
        assert it == 41 and q.final_iters == 41 and q.converged == 1
        assert any(ln.startswith("eigcg_init m 20 nvecs 4 max 10") for ln in log) == (call == 0)
        assert ep.Nvecs_curr == ncurr and ep.Nvecs == nvecs_next
        got = [int(ln.split()[1]) for ln in log if ln.startswith("eigcg_vec_download")]
        assert got == list(range(0 if call == 0 else expect[call - 1][0], ncurr))      # only the new ones
        assert all(vecs[j][0, 0, 0] == 100.0 + j for j in range(ncurr))
        # H[k + ld*j] = H_{k,j}; the stub's row-major entry (k, j) has real part 0.5 * 2 * (k*ld + j)
        assert ep.H[1 + nmax * 2].real == 0.5 * 2 * (1 * nmax + 2) and ep.H[1 + nmax * 2].imag == 0.5 * (2 * (1 * nmax + 2) + 1)
    shim.calc_eigenpairs_gpu(val.ctypes.data, pv, C.byref(ep), EVEN)
    log = _log(shim)
    assert log[0] == "eigcg_pairs 10" and val[3] == 0.004 and ep.H[3 + nmax * 3].real == 0.004 and ep.H[2 + nmax * 3].real == 0.0
    assert len([ln for ln in log if ln.startswith("eigcg_vec_download")]) == 10
    assert shim.b200ks_milc_total_iters() >= 3 * 41


def test_meson_cont_mom_grouping_and_normalisation(shim):
    """ks_meson_cont_mom_gpu around the device contraction (generic_ks/ks_meson_mom.c:160-437): one call per sink
    spin-taste assignment with ITS momenta, local operators as gamma bits, norm_v's phase and factor, accumulation
    into prop[corr_index][t]; a link-shift operator is refused outside a MILC tree (it needs spin_taste_op_fn)."""
    class Cx(C.Structure):
        _fields_ = [("real", C.c_double), ("imag", C.c_double)]
    nt = DIMS[3]
    rng = np.random.default_rng(3)
    s1, s2 = rng.standard_normal((V, 3, 2)), rng.standard_normal((V, 3, 2))
    mom = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 1]], dtype=np.int32)
    par = np.array([[3, 3, 3], [2, 3, 3], [3, 1, 2]], dtype=np.int8)
    pm = (C.POINTER(C.c_int) * 3)(*[mom[k].ctypes.data_as(C.POINTER(C.c_int)) for k in range(3)])
    pp = (C.c_char_p * 3)(*[par[k].tobytes() for k in range(3)])
    # correlators: pion5 (index 0) with momenta 0, 2; rhox0 (13) with momentum 1; G5-G5 (128 + 4*16 + 4) with 0
    spin_taste = [0, 0, 13, 196]
    p_index = [0, 2, 1, 0]
    phase = [0, 1, 2, 3]
    factor = [1.0, 2.0, 0.5, 4.0]
    corr_index = [0, 1, 1, 0]
    groups = [[0, 1], [2], [3]]
    ct = (C.POINTER(C.c_int) * 3)(*[(C.c_int * len(g))(*g) for g in groups])
    nprop = 2
    rows = [(Cx * nt)() for _ in range(nprop)]
    for m in range(nprop):
        for t in range(nt):
            rows[m][t].real, rows[m][t].imag = 0.25 * m, -1.0      # accumulated onto, not overwritten
    prop = (C.POINTER(Cx) * nprop)(*[C.cast(r, C.POINTER(Cx)) for r in rows])
    ia = lambda a: (C.c_int * len(a))(*a)
    shim.ks_meson_cont_mom_gpu.restype = None
    shim.ks_meson_cont_mom_gpu(prop, s1.ctypes.data_as(C.c_void_p), s2.ctypes.data_as(C.c_void_p), 3, pm, pp, 3, ia([2, 1, 1]), ct,
                               ia(p_index), None, None, ia(spin_taste), ia(phase), (C.c_double * 4)(*factor), ia(corr_index),
                               ia([1, 0, 1, 0]))
    log = _log(shim)
    calls = [ln for ln in log if ln.startswith("meson_mom")]
    assert calls == ["meson_mom spin 15 nmom 2 r0 1 0 1 0 same 0 prec 2", "meson_mom spin 9 nmom 1 r0 1 0 1 0 same 0 prec 2",
                     "meson_mom spin 15 nmom 1 r0 1 0 1 0 same 0 prec 2"]
    want = np.zeros((nprop, nt), complex)
    want[0] += 0.0 - 1j
    want[1] += 0.25 - 1j
    ph = [1, 1j, -1, -1j]
    for g, spin in zip(groups, (15, 9, 15)):
        for k, c in enumerate(g):
            p = p_index[c]
            im = mom[p, 0] + 2 * mom[p, 1] + 4 * mom[p, 2] + 0.125 * par[p].sum()
            for t in range(nt):
                want[corr_index[c], t] += ph[phase[c]] * factor[c] * complex(1000.0 * spin + 10.0 * t + k, im)
    got = np.array([[complex(rows[m][t].real, rows[m][t].imag) for t in range(nt)] for m in range(nprop)])
    assert np.allclose(got, want, rtol=0, atol=1e-12)

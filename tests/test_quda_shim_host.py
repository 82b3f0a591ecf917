"""CPU test of the route-2 shim's host logic (milc_qcd_b200/csrc/quda_shim.cu): the quda* entry points compiled
(nvcc, host) against a recording stand-in for the b200ks C ABI (tests/host/b200ks_stub.c).  Checks the argument
marshalling MILC's glue relies on (generic_ks/d_congrad5_fn_gpu.c:95-148, fermion_force_hisq_multi.c:2169-2290):
iteration split, parity mapping, link-cache decisions, and the fermion force's coefficient table with Naik-epsilon
terms and the force filter.  One subprocess per scenario: the shim keeps process-global state like QUDA does."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libquda_shim_stub.so")


@pytest.fixture(scope="module")
def shim_so():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    src = os.path.join(ROOT, "milc_qcd_b200", "csrc", "quda_shim.cu")
    stub = os.path.join(HOST_DIR, "b200ks_stub.c")
    deps = [src, stub] + [os.path.join(ROOT, "include", f) for f in ("b200ks.h", "quda_milc_interface.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in deps):
        obj = os.path.join(HOST_DIR, "b200ks_stub.o")
        subprocess.check_call(["gcc", "-O1", "-fPIC", "-std=gnu99", "-I", os.path.join(ROOT, "include"), "-c", stub, "-o", obj])
        subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-x", "cu", "--shared", "-Xcompiler", "-fPIC",
                               "-Wno-deprecated-gpu-targets", "-Xlinker", "-Bsymbolic", "-o", SO, src, "-Xlinker", obj])
    return SO


PRELUDE = r"""
import ctypes as C, numpy as np
lib = C.CDLL(%r)
lib.stub_log.restype = C.c_char_p
class Layout(C.Structure):
    _fields_ = [("latsize", C.POINTER(C.c_int)), ("machsize", C.POINTER(C.c_int)), ("device", C.c_int)]
class InitArgs(C.Structure):
    _fields_ = [("verbosity", C.c_int), ("layout", Layout)]
class InvertArgs(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("evenodd", C.c_int), ("mixed_precision", C.c_int), ("boundary_phase", C.c_double * 4),
                ("tadpole", C.c_double), ("naik_epsilon", C.c_double)]
class HisqParams(C.Structure):
    _fields_ = [("reunit_allow_svd", C.c_int), ("reunit_svd_only", C.c_int), ("reunit_svd_abs_error", C.c_double),
                ("reunit_svd_rel_error", C.c_double), ("force_filter", C.c_double)]
dims = (C.c_int * 4)(4, 4, 2, 2)
mach = (C.c_int * 4)(1, 1, 1, 1)
V = 64
lib.qudaInit.argtypes = [InitArgs]
lib.qudaInit(InitArgs(1, Layout(dims, mach, 0)))
def log():
    out = lib.stub_log().decode(); lib.stub_reset()
    return [ln for ln in out.splitlines() if ln]
"""


def _run(shim_so, body):
    out = subprocess.run([sys.executable, "-c", PRELUDE % shim_so + body], capture_output=True, text=True, timeout=300)
    assert "HOST-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_invert_marshalling_and_link_cache(shim_so):
    _run(shim_so, r"""
rng = np.random.default_rng(1)
fat, lng = rng.standard_normal((V, 4, 18)), rng.standard_normal((V, 4, 18))
src, sol = rng.standard_normal((V, 6)), np.zeros((V, 6))
lib.qudaInvert.argtypes = [C.c_int, C.c_int, C.c_double, InvertArgs, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
res, rel, it = C.c_double(), C.c_double(), C.c_int(-1)
args = InvertArgs(1500, 1, 2)          # qic->max * qic->nrestart, QUDA_ODD_PARITY, MAX_MIXED
def call():
    lib.qudaInvert(2, 2, 0.05, args, 1e-9, 0.0, fat.ctypes.data, lng.ctypes.data, src.ctypes.data, sol.ctypes.data,
                   C.byref(res), C.byref(rel), C.byref(it))
    return log()
l = call()
assert l[0].startswith('create 4 4 2 2') and l[1].startswith('load_links prec 2'), l
assert 'congrad mass 0.05 parity 1 max 300 nrestart 5 resid 1e-09 relresid 0 mixed 2 prec 2' in l[2], l
assert it.value == 17 and abs(res.value - 1e-10) < 1e-24
assert np.array_equal(sol[V // 2:], src[V // 2:]) and np.all(sol[:V // 2] == 0)
assert not any(x.startswith('load_links') for x in call())          # nothing changed: no upload
lng[5, 2, 7] += 0.5                                                  # in-place edit without notice
assert any(x.startswith('load_links') for x in call())
it.value = -1                                                        # MILC's "links changed" signal
assert any(x.startswith('load_links') for x in call())
print('HOST-OK')
""")


def test_hisq_force_coefficients_naik_terms_and_filter(shim_so):
    _run(shim_so, r"""
nterms, nnaik = 3, 1
coeff = [(C.c_double * 2)(2 * r, -2 * r / 24) for r in (0.5, 0.25, 0.125)] + [(C.c_double * 2)(0.01, -0.02)]
cp = (C.POINTER(C.c_double) * (nterms + nnaik))(*[C.cast(c, C.POINTER(C.c_double)) for c in coeff])
xs = [np.full((V, 6), 10.0 + j) for j in range(nterms)]
xp = (C.c_void_p * nterms)(*[x.ctypes.data for x in xs])
l2 = (C.c_double * 6)(1.0, -1 / 24, -1 / 16, 1 / 64, -1 / 384, -1 / 8)
f7 = (C.c_double * 6)(1 / 8, 0, -1 / 16, 1 / 64, -1 / 384, 0)
W, Vl, U, mom = (np.zeros((V, 4, 18)) for _ in range(4))
lib.qudaHisqForce.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_void_p),
                              C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
def call():
    lib.qudaHisqForce(2, nterms, nnaik, 0.02, cp, xp, l2, f7, W.ctypes.data, Vl.ctypes.data, U.ctypes.data, mom.ctypes.data)
    return [x for x in log() if x.startswith('hisq_force')]
l = call()       # before qudaHisqParamsInit: no filter
assert l == ['hisq_force nterms 3 naik 1 eps 0.02 filter 0 prec 2 coeff 1 -0.0416667 0.5 -0.0208333 0.25 -0.0104167 0.01 -0.02 '
             'x0 10 11 12 l2 1 -0.125 f7 0.125 -0.0625'], l
lib.qudaHisqParamsInit.argtypes = [HisqParams]
lib.qudaHisqParamsInit(HisqParams(1, 0, 1e-8, 1e-8, 5e-5))
l = call()
assert 'filter 5e-05' in l[0], l
print('HOST-OK')
""")


def test_multishift_block_dslash_and_link_construction_marshalling(shim_so):
    _run(shim_so, r"""
class FatLinkArgs(C.Structure):
    _fields_ = [("su3_source", C.c_int)]
rng = np.random.default_rng(2)
fat, lng = rng.standard_normal((V, 4, 18)), rng.standard_normal((V, 4, 18))
src = rng.standard_normal((V, 6))
# multi-shift: the cap is the product (no restarts), convergence on target_residual[0], one residual per shift
n = 3
off = (C.c_double * n)(0.01, 0.04, 0.25)
tr = (C.c_double * n)(1e-6, 1e-5, 1e-4)
trf = (C.c_double * n)(0, 0, 0)
sols = [np.zeros((V, 6)) for _ in range(n)]
sp = (C.c_void_p * n)(*[s.ctypes.data for s in sols])
fr, ffr, it = (C.c_double * n)(), (C.c_double * n)(), C.c_int(-1)
lib.qudaMultishiftInvert.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), InvertArgs, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
lib.qudaMultishiftInvert(2, 2, n, off, InvertArgs(2500, 0, 0), tr, trf, fat.ctypes.data, lng.ctypes.data, src.ctypes.data, sp,
                         fr, ffr, C.byref(it))
l = log()
assert any(x == 'multicg n 3 parity 2 offsets 0.01 0.04 0.25' for x in l), l
assert it.value == 23 and all(abs(fr[j] - 1e-10) < 1e-24 and ffr[j] == 0 for j in range(n))
assert all(np.array_equal(s[:V // 2], src[:V // 2]) for s in sols)
# block solve: every source to b200ks_congrad_block, worst residual and the total iteration count back
ns = 3
srcs = [rng.standard_normal((V, 6)) for _ in range(ns)]
dsts = [np.zeros((V, 6)) for _ in range(ns)]
sa = (C.c_void_p * ns)(*[s.ctypes.data for s in srcs])
da = (C.c_void_p * ns)(*[d.ctypes.data for d in dsts])
res, rel = C.c_double(), C.c_double()
lib.qudaInvertMsrc.argtypes = [C.c_int, C.c_int, C.c_double, InvertArgs, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                               C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.POINTER(C.c_double),
                               C.POINTER(C.c_int), C.c_int]
lib.qudaInvertMsrc(2, 2, 0.1, InvertArgs(1000, 0, 0), 1e-8, 0.0, fat.ctypes.data, lng.ctypes.data, sa, da, C.byref(res),
                   C.byref(rel), C.byref(it), ns)
l = log()
assert any(x == 'congrad_block nsrc 3 mass 0.1 parity 2' for x in l) and not any(x.startswith('load_links') for x in l), l
assert it.value == 33 and abs(res.value - 1e-10) < 1e-24
# dslash
out = np.zeros((V, 6))
lib.qudaDslash.argtypes = [C.c_int, C.c_int, InvertArgs, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
lib.qudaDslash(2, 2, InvertArgs(0, 1, 0), fat.ctypes.data, lng.ctypes.data, src.ctypes.data, out.ctypes.data, C.byref(it))
assert 'dslash parity 1' in log() and it.value == 0
# link construction: coefficients and the optional outputs pass through
coef = (C.c_double * 6)(1.0, -1 / 24, -1 / 16, 1 / 64, -1 / 384, -1 / 8)
U, F, L = (np.zeros((V, 4, 18)) for _ in range(3))
lib.qudaLoadKSLink.argtypes = [C.c_int, FatLinkArgs, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]
lib.qudaLoadKSLink(2, FatLinkArgs(1), coef, U.ctypes.data, F.ctypes.data, L.ctypes.data)
lib.qudaLoadKSLink(1, FatLinkArgs(1), coef, U.ctypes.data, F.ctypes.data, None)
lib.qudaLoadUnitarizedLink.argtypes = lib.qudaLoadKSLink.argtypes
lib.qudaLoadUnitarizedLink(2, FatLinkArgs(1), coef, U.ctypes.data, None, L.ctypes.data)
assert log() == ['ks_links c0 1 naik -0.0416667 long 1 prec 2', 'ks_links c0 1 naik -0.0416667 long 0 prec 1',
                 'unitarized_links c0 1 v 0 prec 2']
print('HOST-OK')
""")

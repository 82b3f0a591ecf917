"""CPU test of the CUDA deflation's arithmetic: the __host__ __device__ site
routines of milc_qcd_b200/csrc/deflate.cuh run in host loops with the kernels' structure
(tests/host/deflate_host.cu, compiled with nvcc for the host) against the CPU oracle (oracle/ks_oracle.c
kso_deflate, pinned on the reference's deflated mat_invert_uml_field)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libdeflate_host.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host_deflate():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    src = os.path.join(HOST_DIR, "deflate_host.cu")
    hdrs = [os.path.join(ROOT, "milc_qcd_b200", "csrc", f) for f in ("deflate.cuh", "common.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-x", "cu", "--shared", "-Xcompiler", "-fPIC", "-o", SO, src])
    lib = C.CDLL(SO)
    lib.deflate_host.restype = None
    lib.deflate_host.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int]
    return lib


@pytest.mark.parametrize("nvecs,nchunks", [(1, 1), (16, 7), (48, 296), (8, 1184)])
def test_deflate_site_routines_match_oracle(host_deflate, nvecs, nchunks):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_deflate import low_modes
    from milc_qcd_b200 import fields as F
    from oracle.pyoracle import Oracle, EVEN, ODD, EVENANDODD
    dims = (4, 4, 4, 4)
    V = 256
    fat, lng = F.make_links(dims, seed=77)
    o = Oracle()
    lam, ev = low_modes(o, dims, fat, lng)
    ev, lam = np.ascontiguousarray(ev[:nvecs]), np.ascontiguousarray(lam[:nvecs])
    src = F.make_source(dims, seed=31, parity=EVENANDODD)
    rng = np.random.default_rng(6)
    for parity, pbit in ((EVEN, 0), (ODD, 1)):
        guess = rng.standard_normal(src.shape)
        want = o.deflate(dims, guess.copy(), src, 0.03, ev, lam, parity)
        got = guess.copy()
        host_deflate.deflate_host(V, nvecs, ev, lam, src, got, 0.03, pbit, nchunks)
        h = V // 2
        other = slice(h, V) if pbit == 0 else slice(0, h)
        assert np.array_equal(got[other], guess[other])            # only the sites of `parity` are touched
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        assert np.abs(got - guess).max() > 1e-3                     # and it did something

"""CPU test of the eigCG solver's host-side dense algebra (milc_qcd_b200/csrc/eigcg.cuh namespace dense: Jacobi
Hermitian eigensolver, Gram-Schmidt, Cholesky solve -- what the reference takes from LAPACK, generic_ks/inc_eigcg.c)
against numpy."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "tests", "host")
SO = os.path.join(HOST_DIR, "libeigcg_host.so")


@pytest.fixture(scope="module")
def lib():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    src = os.path.join(HOST_DIR, "eigcg_host.cu")
    deps = [src, os.path.join(ROOT, "milc_qcd_b200", "csrc", "eigcg.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-o", SO, src])
    return C.CDLL(SO)


def _cplx(a):
    return np.ascontiguousarray(np.stack([a.real, a.imag], axis=-1))


@pytest.mark.parametrize("n", [1, 2, 7, 24, 60])
def test_jacobi_heev_matches_numpy(lib, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.conj().T
    if n >= 7:   # a cluster of nearly degenerate values and a tridiagonal-plus-arrow shape like the Lanczos matrix
        A = np.diag(np.concatenate([np.full(3, 0.5) + 1e-9 * np.arange(3), rng.uniform(1, 2, n - 3)])).astype(complex)
        A[:4, 4] = rng.standard_normal(4) + 1j * rng.standard_normal(4)
        for k in range(4, n - 1):
            A[k, k + 1] = rng.standard_normal()
    Aup = np.triu(A)      # only the upper triangle is read
    Ah = np.triu(A) + np.triu(A, 1).conj().T
    w, Z = np.zeros(n), np.zeros((n, n, 2))
    lib.host_heev(n, _cplx(Aup).ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), Z.ctypes.data_as(C.c_void_p))
    Zc = Z[..., 0] + 1j * Z[..., 1]
    wn = np.linalg.eigvalsh(Ah)
    scale = max(np.abs(wn).max(), 1.0)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - wn).max() <= 1e-12 * scale
    assert np.abs(Zc.conj().T @ Zc - np.eye(n)).max() <= 1e-12
    assert np.abs(Ah @ Zc - Zc * w).max() <= 1e-11 * scale


def test_orthonormalize_and_posv(lib):
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((40, 12)) + 1j * rng.standard_normal((40, 12))
    Y[:, 7] = Y[:, 2] + 1e-9 * Y[:, 7]          # nearly dependent column
    Yr = _cplx(Y)
    lib.host_orthonormalize(40, 12, Yr.ctypes.data_as(C.c_void_p))
    Q = Yr[..., 0] + 1j * Yr[..., 1]
    assert np.abs(Q.conj().T @ Q - np.eye(12)).max() <= 1e-10
    # the span is that of Y
    P = Q @ Q.conj().T
    assert np.abs(P @ Y - Y).max() <= 1e-9 * np.abs(Y).max()
    n = 17
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = B @ B.conj().T + 0.01 * np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = _cplx(b)
    assert lib.host_posv(n, _cplx(np.triu(A)).ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)) == 0
    xc = x[..., 0] + 1j * x[..., 1]
    assert np.abs(A @ xc - b).max() <= 1e-10 * np.abs(b).max()
    bad = _cplx(np.triu(-A))
    assert lib.host_posv(n, bad.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)) == -1

"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/b200ks.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ks_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from milc_qcd_b200 import _lib, build
    build.build_all()
    lib = _lib.load()
    declared = _declared("b200ks.h")
    assert len(declared) >= 20
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200ks_version() == 121


def test_no_cpu_fallback_without_gpu():
    import torch
    from milc_qcd_b200 import _lib
    lib = _lib.load()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.b200ks_device_count() == 0
    dims = (C.c_int * 4)(4, 4, 4, 4)
    assert not lib.b200ks_create(dims, 0)
    assert b"no CPU fallback" in lib.b200ks_last_error()
    from milc_qcd_b200 import api
    with pytest.raises(_lib.B200KSError):
        api.Context((4, 4, 4, 4))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or the reference build)."""
    pkg = os.path.join(ROOT, "milc_qcd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "libmilcref" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_fingerprint_sees_negated_time_slices():
    """boundary_twist_fn (generic_ks/fermion_links_fn_twist_milc.c:318-400) negates whole time slices of
    links in place.  A hash that is linear mod 2^64 cannot see an even number of sign flips per lane
    (round-1 defect); the real b200ks_fingerprint must."""
    import numpy as np
    from milc_qcd_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.float32):
        dims = (8, 8, 8, 8)
        V = int(np.prod(dims))
        links = rng.standard_normal((V, 4, 3, 3, 2)).astype(dtype)
        fp = lambda a: lib.b200ks_fingerprint(a.ctypes.data_as(C.c_void_p), a.nbytes)
        f0 = fp(links)
        assert f0 == fp(links.copy())
        # even sites then odd sites, index inside a parity block = lex/2, t slowest: the last time slice
        # of each parity block is its last V/(2*8) sites
        sl = V // 2 // dims[3]
        twisted = links.copy()
        for blk in (0, V // 2):
            twisted[blk + V // 2 - sl: blk + V // 2, 3] *= -1      # all t links of the last slice
        assert fp(twisted) != f0
        three = links.copy()
        for blk in (0, V // 2):
            three[blk + V // 2 - 3 * sl: blk + V // 2] *= -1       # every link of the last three slices (Naik)
        assert fp(three) != f0 and fp(three) != fp(twisted)
        one = links.copy()
        one.reshape(-1)[[5, 5 + 8 * 2]] *= -1                      # two flips in the same hash lane
        assert fp(one) != f0
        twisted_back = twisted.copy()
        for blk in (0, V // 2):
            twisted_back[blk + V // 2 - sl: blk + V // 2, 3] *= -1
        assert fp(twisted_back) == f0

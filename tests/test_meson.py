"""Meson tie-ups (generic_ks/ks_meson_mom.c:160-437).

CPU: the oracle's contraction (oracle/ks_oracle.c kso_meson_mom) inside the host-side mirror of the reference interface
(milc_qcd_b200/meson.py ks_meson_cont_mom) is pinned on the committed golden tests/golden/ref_meson.npz, produced by the
reference's own ks_meson_cont_mom (tests/golden/make_golden_meson.py), and live on oracle/_ref when it is there.
GPU: the library's contraction (b200ks_meson_mom[_dev], csrc/meson.cuh) through the same mirror against the golden, and
against the oracle on more lattices, origins and momentum sets; host precision; device-resident propagators."""
import os

import numpy as np
import pytest

from conftest import ROOT

import meson_case as K
from milc_qcd_b200 import meson as M

GOLD = os.path.join(ROOT, "tests", "golden", "ref_meson.npz")


def _gold():
    g = np.load(GOLD)
    index_of = {str(n): int(i) for n, i in zip(g["names"], g["index"])}
    ops = {str(k): f for k, f in zip(g["op_keys"], g["op_fields"])}
    return g, index_of, ops


def _golden_op(index_of, ops, s1, s2):
    """spin_taste_op_fn of the reference, from the stored fields (keyed by operator index and which propagator)."""
    name_of = {v: k for k, v in index_of.items()}

    def op(index, r0, field):
        for nm, i in index_of.items():
            if M.is_rhosfn(i) and index == M.backward_index(i) and field is s1:
                return ops[nm + "/b1"]
            if M.is_rhosfn(i) and index == M.forward_index(i) and field is s2:
                return ops[nm + "/f2"]
        assert field is s1
        return ops[name_of[index] + "/1"]
    return op


def _run(contract, names, index_of, s1, s2, op):
    st, pi, ph, fa, ci, ct = K.table(index_of, names)
    prop = np.zeros((K.NPROP, K.DIMS[3]), complex)
    return M.ks_meson_cont_mom(contract, prop, s1, s2, len(K.MOM), K.MOM, K.PAR, len(ct), [len(t) for t in ct], ct, pi, st, ph, fa,
                               ci, K.R0, spin_taste_op=op)


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_local_operator_table_matches_the_reference_indices():
    g, index_of, ops = _gold()
    for nm in K.LOCAL:
        assert M.local_spin_bits(index_of[nm]) is not None, nm
    for nm in K.SHIFTED:
        assert M.local_spin_bits(index_of[nm]) is None, nm
    assert [M.local_spin_bits(index_of[n]) for n in ("pion5", "pion05", "rhox", "rhoz0", "GT-GT")] == [15, 0, 1, 12, 8]
    for nm in ("pion5", "pion05", "rhoi", "rhox", "rhoz0", "rhoxsfn", "rhotsfn", "rhozs"):
        assert M.INDEX[nm] == index_of[nm]


def test_meson_oracle_reproduces_the_reference_golden(oracle):
    g, index_of, ops = _gold()
    s1, s2 = K.sources()
    contract = lambda a, q, spin, r0, mom, par: oracle.meson_mom(K.DIMS, a, q, spin, r0, mom, par)   # noqa: E731
    op = _golden_op(index_of, ops, s1, s2)
    assert _rel(_run(contract, K.LOCAL, index_of, s1, s2, op), g["prop_local"]) <= 1e-13
    assert _rel(_run(contract, K.SHIFTED, index_of, s1, s2, op), g["prop_shifted"]) <= 1e-13


def test_meson_oracle_matches_the_compiled_reference_live(oracle):
    from oracle import pyoracle
    from milc_qcd_b200 import fields as F
    if not pyoracle.ref_available(""):
        pytest.skip("oracle/_ref not built")
    dims = (4, 4, 4, 8)      # (one MILC geometry per process: the one tests/test_eigcg.py uses)
    ref = pyoracle.MilcRef(dims, "")
    if not ref.has_meson:
        pytest.skip("oracle/_ref was built without the meson harness")
    fat, lng = F.make_links(dims, seed=3)
    ref.set_links(fat, lng)
    ref.set_ape_links(F.make_links(dims, seed=4)[0])
    rng = np.random.default_rng(8)
    V = int(np.prod(dims))
    s1, s2 = rng.standard_normal((V, 3, 2)), rng.standard_normal((V, 3, 2))
    names = ["pion05", "rhoy0", "GYZ-GYZ", "rhoysfn", "pions"]
    index_of = {nm: ref.spin_taste_index(nm) for nm in names}
    st, pi, ph, fa, ci, ct = K.table(index_of, names)
    r0 = [0, 3, 1, 2]
    want = ref.meson_cont_mom(s1, s2, K.MOM, K.PAR, st, pi, ph, fa, ci, K.NPROP, r0)
    contract = lambda a, q, spin, r0, mom, par: oracle.meson_mom(dims, a, q, spin, r0, mom, par)   # noqa: E731
    prop = np.zeros((K.NPROP, dims[3]), complex)
    got = M.ks_meson_cont_mom(contract, prop, s1, s2, len(K.MOM), K.MOM, K.PAR, len(ct), [len(t) for t in ct], ct, pi, st, ph, fa, ci,
                              r0, spin_taste_op=ref.spin_taste_op)
    assert _rel(got, want) <= 1e-13


@pytest.mark.gpu
def test_meson_tieups_on_the_gpu_match_the_reference_golden():
    from milc_qcd_b200 import api
    g, index_of, ops = _gold()
    s1, s2 = K.sources()
    ctx = api.Context(K.DIMS)
    op = _golden_op(index_of, ops, s1, s2)
    assert _rel(_run(ctx.meson_mom, K.LOCAL, index_of, s1, s2, op), g["prop_local"]) <= 1e-13
    assert _rel(_run(ctx.meson_mom, K.SHIFTED, index_of, s1, s2, op), g["prop_shifted"]) <= 1e-13
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dims,r0", [((8, 8, 8, 8), (0, 0, 0, 0)), ((4, 4, 4, 4), (1, 2, 3, 0)), ((16, 12, 8, 6), (5, 0, 7, 3)),
                                     ((2, 2, 2, 2), (0, 1, 0, 1)), ((32, 32, 8, 4), (0, 0, 0, 0))])
def test_meson_contraction_matches_oracle(oracle, dims, r0):
    """More than one chunk per time slice (32 x 32 x 8: 4096 sites per slice and parity), extent-2 lattices, every local
    operator, 23 momenta in one call, single-precision host fields, antiquark == quark, resident propagators."""
    from milc_qcd_b200 import api
    rng = np.random.default_rng(17)
    V = int(np.prod(dims))
    s1, s2 = rng.standard_normal((V, 3, 2)), rng.standard_normal((V, 3, 2))
    mom = rng.integers(-2, 3, size=(23, 3))
    par = rng.integers(1, 4, size=(23, 3))
    ctx = api.Context(dims)
    for spin in (-1, 0, 5, 9, 15):
        want = oracle.meson_mom(dims, s1, s2, spin, r0, mom, par)
        got = ctx.meson_mom(s1, s2, spin, r0, mom, par)
        assert got.shape == (dims[3], 23)
        assert _rel(got, want) <= 1e-13
    # norm of one field: pion5 with itself at zero momentum is sum |s|^2 per time slice (real, positive)
    got = ctx.meson_mom(s1, s1, 15, r0, [[0, 0, 0]], [[3, 3, 3]])
    assert np.all(got.real > 0) and np.abs(got.imag).max() <= 1e-12 * got.real.max()
    assert abs(got.real.sum() - np.sum(s1 * s1)) <= 1e-12 * np.sum(s1 * s1)
    # float host fields (a PRECISION=1 MILC build)
    f1, f2 = s1.astype(np.float32), s2.astype(np.float32)
    want = oracle.meson_mom(dims, f1.astype(np.float64), f2.astype(np.float64), 9, r0, mom, par)
    assert _rel(ctx.meson_mom(f1, f2, 9, r0, mom, par), want) <= 1e-13
    # propagators that never left the device
    va, vq = ctx.vec_create(), ctx.vec_create()
    ctx.vec_upload(va, s1)
    ctx.vec_upload(vq, s2)
    want = oracle.meson_mom(dims, s1, s2, 5, r0, mom, par)
    assert _rel(ctx.meson_mom_dev(va, vq, 5, r0, mom, par), want) <= 1e-13
    ctx.close()


@pytest.mark.gpu
def test_meson_contraction_refuses_bad_arguments():
    from milc_qcd_b200 import api, _lib
    ctx = api.Context((4, 4, 4, 4))
    s = np.zeros((256, 3, 2))
    with pytest.raises(_lib.B200KSError):
        ctx.meson_mom(s, s, 16, (0, 0, 0, 0), [[0, 0, 0]], [[3, 3, 3]])          # not a gamma bit pattern
    with pytest.raises(_lib.B200KSError):
        ctx.meson_mom(s, s, 15, (0, 0, 0, 0), [[0, 0, 0]], [[3, 0, 3]])          # not a parity code
    with pytest.raises(_lib.B200KSError):
        ctx.meson_mom(s, s, 15, (0, 0, 0, 0), np.zeros((129, 3), int), np.full((129, 3), 3))   # more than 128 momenta
    ctx.close()

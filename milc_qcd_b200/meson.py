"""Host-side mirror of the reference's meson tie-up interface.

`ks_meson_cont_mom` has the argument list of generic_ks/ks_meson_mom.c:160-178 and does what the route-1 shim
(csrc_milc/milc_shim.c ks_meson_cont_mom_gpu) does around the device contraction b200ks_meson_mom:

* one device call per sink spin-taste assignment g (corr_table[g]), with the momenta that assignment uses;
* LOCAL sink operators (site signs: pion5, pion05, rhox/y/z, rhox0/y0/z0 and every gamma-gamma operator whose spin and
  taste agree; generic_ks/spin_taste_ops.c:172-263) are applied inside the kernel (`spin` = gamma bits); every other
  operator is built from link shifts by MILC's own spin_taste_op_fn on the host (`spin_taste_op` argument) and the
  contraction runs with spin = -1;
* the quasi-conserved vector currents average the backward operator on the antiquark and the forward operator on the
  quark (ks_meson_mom.c:296-312, 349-356);
* correlator phase, factor and accumulation into prop[corr_index[c]][t] as norm_v and the loop behind it
  (ks_meson_mom.c:100-131, 405-419).

The tests call it with the library's contraction (api.Context.meson_mom) and with the CPU oracle's, against the
reference's own ks_meson_cont_mom.
"""
import numpy as np

# enum spin_taste_type, generic_ks/spin_taste_ops.c:805-853
SPIN_TASTE = ["pion5", "pion05", "pioni5", "pionij", "pioni", "pioni0", "pions", "pion0", "rhoi", "rhox", "rhoy", "rhoz",
              "rhoi0", "rhox0", "rhoy0", "rhoz0", "rhoxs", "rhoys", "rhozs", "rhots", "rhois", "rho0",
              "rhoxsfn", "rhoysfn", "rhozsfn", "rhotsfn", "rhoxsffn", "rhoysffn", "rhozsffn", "rhotsffn",
              "rhoxsbfn", "rhoysbfn", "rhozsbfn", "rhotsbfn", "rhoxsape", "rhoysape", "rhozsape", "rhotsape",
              "rhoxsfape", "rhoysfape", "rhozsfape", "rhotsfape", "rhoxsbape", "rhoysbape", "rhozsbape", "rhotsbape"]
INDEX = {name: k for k, name in enumerate(SPIN_TASTE)}
# gamma_hex_value, generic_wilson/gammas.c:21-22 (enum gammatype order GX GY GZ GT G5 GYZ GZX GXY GXT GYT GZT G5X G5Y G5Z G5T G1)
GAMMA_HEX = [1, 2, 4, 8, 15, 6, 5, 3, 9, 10, 12, 14, 13, 11, 7, 0]
_LOCAL = {"pion5": 15, "pion05": 0, "rhox": 1, "rhoy": 2, "rhoz": 4, "rhoi": 4, "rhox0": 9, "rhoy0": 10, "rhoz0": 12, "rhoi0": 12}


def local_spin_bits(index):
    """gamma bits of a LOCAL sink operator, or None when the operator needs link shifts."""
    if index >= 128:                                   # gamma-gamma style, spin_taste_ops.c:960-973
        s, t = (index - 128) // 16, (index - 128) % 16
        return GAMMA_HEX[s] if s == t else None
    return _LOCAL.get(SPIN_TASTE[index])


def _family(index, first):
    return index < 128 and INDEX[first] <= index < INDEX[first] + 4


def is_rhosfn(i): return _family(i, "rhoxsfn") or _family(i, "rhoxsape")      # noqa: E704  spin_taste_ops.c:983-1019
def is_rhosffn(i): return _family(i, "rhoxsffn") or _family(i, "rhoxsfape")   # noqa: E704
def is_rhosbfn(i): return _family(i, "rhoxsbfn") or _family(i, "rhoxsbape")   # noqa: E704
def forward_index(i): return i + 4 if is_rhosfn(i) else -1                    # noqa: E704  :1041-1064 (-1 as there)
def backward_index(i): return i + 8 if is_rhosfn(i) else -1                   # noqa: E704  :1067-1090


_PHASE = [1.0, 1j, -1.0, -1j]          # meson_phase encoding, include/gammatypes.h


def ks_meson_cont_mom(contract, prop, src1, src2, no_q_momenta, q_momstore, q_parity, no_spin_taste_corr, num_corr_mom,
                      corr_table, p_index, spin_taste_snk, meson_phase, meson_factor, corr_index, r0, spin_taste_op=None):
    """prop[m][t] += ... like the reference.  contract(antiquark, quark, spin, r0, mom, mom_parity) -> corr[t][k] is the
    device contraction (api.Context.meson_mom) or the oracle's; spin_taste_op(index, r0, field) -> field applies a
    non-local sink operator (MILC's spin_taste_op_fn)."""
    assert no_q_momenta <= 100, "MAXQ"
    q_momstore = np.asarray(q_momstore, dtype=np.int32).reshape(-1, 3)
    q_parity = np.asarray(q_parity, dtype=np.int8).reshape(-1, 3)
    for g in range(no_spin_taste_corr):
        cs = list(corr_table[g][: num_corr_mom[g]])
        st = spin_taste_snk[cs[0]]
        ps = [p_index[c] for c in cs]
        mom, par = q_momstore[ps], q_parity[ps]
        bits = local_spin_bits(st)
        if bits is not None:
            corr = contract(src1, src2, bits, r0, mom, par)
        elif is_rhosfn(st):
            corr = 0.5 * (contract(spin_taste_op(backward_index(st), r0, src1), src2, -1, r0, mom, par)
                          + contract(src1, spin_taste_op(forward_index(st), r0, src2), -1, r0, mom, par))
        elif is_rhosffn(st):
            corr = contract(src1, spin_taste_op(forward_index(st), r0, src2), -1, r0, mom, par)
        elif is_rhosbfn(st):
            corr = contract(spin_taste_op(backward_index(st), r0, src1), src2, -1, r0, mom, par)
        else:   # every other link-shift operator: on the antiquark
            corr = contract(spin_taste_op(st, r0, src1), src2, -1, r0, mom, par)
        for k, c in enumerate(cs):
            prop[corr_index[c], :] += _PHASE[meson_phase[c]] * meson_factor[c] * corr[:, k]
    return prop

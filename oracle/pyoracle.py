"""oracle/pyoracle.py -- TEST INFRASTRUCTURE (ctypes loaders for the checkers).

Two checkers live behind this module:

* ``Oracle``   -- our C restatement ``oracle/ks_oracle.c`` (always buildable).
* ``MilcRef``  -- the reference's own CPU sources compiled by ``oracle/build_ref.sh``
  into ``oracle/_ref/libmilcref*.so`` (present when built in the container; the
  prebuilt .so travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  The product (milc_qcd_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EVEN, ODD, EVENANDODD = 2, 1, 3

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("ks_oracle.c", "ks_links_oracle.c", "ks_force_oracle.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liboracle.so"])
    return so


class Oracle:
    """ctypes face of ks_oracle.c.  Arrays are MILC host layout, float64."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.kso_node_index.restype = C.c_long
        L.kso_node_index.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.kso_dslash.restype = None
        L.kso_dslash.argtypes = [_ip, _dp, _dp, _dp, _dp, C.c_int]
        L.kso_relative_residue.restype = C.c_double
        L.kso_relative_residue.argtypes = [_ip, _dp, _dp, C.c_int]
        L.kso_congrad.restype = C.c_int
        L.kso_congrad.argtypes = [_ip, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_int, _dp]
        L.kso_multicg.restype = C.c_int
        L.kso_deflate.restype = None
        L.kso_deflate.argtypes = [_ip, _dp, _dp, C.c_double, C.c_int, _dp, _dp, C.c_int]
        L.kso_multicg.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, _dp]
        L.kso_meson_mom.restype = None
        L.kso_meson_mom.argtypes = [_ip, _dp, _dp, C.c_int, _ip, C.c_int, _ip, C.c_char_p, _dp]

    @staticmethod
    def _dims(dims):
        return np.ascontiguousarray(dims, dtype=np.int32)

    def node_index(self, dims, x, y, z, t):
        return self.lib.kso_node_index(self._dims(dims), x, y, z, t)

    def dslash(self, dims, fat, lng, src, parity, dest=None):
        if dest is None:
            dest = np.zeros_like(src)
        self.lib.kso_dslash(self._dims(dims), fat, lng, src, dest, parity)
        return dest

    def congrad(self, dims, fat, lng, src, dest, mass, parity, niter, nrestart, resid,
                relresid=0.0, fewsums=True):
        out = np.zeros(7)
        it = self.lib.kso_congrad(self._dims(dims), fat, lng, src, dest, mass, parity, niter,
                                  nrestart, resid, relresid, int(fewsums), out)
        return it, _qic(out)

    def deflate(self, dims, dst, src, mass, eigvec, eigval, parity):
        """kso_deflate (deflate() + project_out(), generic_ks/mat_invert.c:131-183): dst updated in place on
        the sites of `parity`; eigvec (nvecs, V, 3, 2) with both parities filled, eigval of -D_eo D_oe."""
        ev = np.ascontiguousarray(eigvec, np.float64)
        self.lib.kso_deflate(self._dims(dims), dst, np.ascontiguousarray(src, np.float64), mass, ev.shape[0], ev,
                             np.ascontiguousarray(eigval, np.float64), parity)
        return dst

    def meson_mom(self, dims, antiquark, quark, spin, r0, mom, mom_parity):
        """kso_meson_mom (site loops of ks_meson_cont_mom, generic_ks/ks_meson_mom.c:160-437, one sink spin-taste
        assignment): corr[t][p] complex.  spin = gamma bits of a local sink operator, -1 = none."""
        mom = np.ascontiguousarray(mom, dtype=np.int32).reshape(-1, 3)
        par = np.ascontiguousarray(mom_parity, dtype=np.int8).reshape(-1, 3)
        out = np.zeros((int(dims[3]), mom.shape[0], 2))
        self.lib.kso_meson_mom(self._dims(dims), np.ascontiguousarray(antiquark, np.float64),
                               np.ascontiguousarray(quark, np.float64), int(spin), np.ascontiguousarray(r0, dtype=np.int32),
                               mom.shape[0], mom, par.tobytes(), out)
        return out[..., 0] + 1j * out[..., 1]

    def multicg(self, dims, fat, lng, src, offsets, parity, niter, nrestart, resid, relresid=0.0):
        offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        n = len(offsets)
        psim = np.zeros((n,) + src.shape)
        out = np.zeros(7 * n)
        it = self.lib.kso_multicg(self._dims(dims), fat, lng, src, psim, offsets, n, parity, niter,
                                  nrestart, resid, relresid, out)
        return it, psim, [_qic(out[7 * j:7 * j + 7]) for j in range(n)]


class LinksOracle:
    """ctypes face of ks_links_oracle.c (HISQ/asqtad link construction).  MILC host layout, float64."""
    # the reference's HISQ coefficients (generic_ks/imp_actions/hisq/hisq_u3_action.h:33-37,74-81):
    # {one_link, naik, three_staple, five_staple, seven_staple, lepage}
    FAT7 = (1.0 / 8.0, 0.0, -1.0 / 16.0, 1.0 / 64.0, -1.0 / 384.0, 0.0)
    ASQTAD_LIKE = (1.0, -1.0 / 24.0, -1.0 / 16.0, 1.0 / 64.0, -1.0 / 384.0, -1.0 / 8.0)

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.ksl_smear.restype = None
        L.ksl_smear.argtypes = [_ip, _dp, _dp, _dp, C.c_void_p]
        L.ksl_unitarize.restype = C.c_long
        L.ksl_unitarize.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_double, C.c_double]
        L.ksf_hisq_force.restype = None
        L.ksf_hisq_force.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, _dp]
        L.ksf_hisq_force_naik.restype = None
        L.ksf_hisq_force_naik.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, _ip, _dp, C.c_double, _dp]
        L.ksf_set_force_filter.restype = None
        L.ksf_set_force_filter.argtypes = [C.c_double]
        L.ksf_unitarize_bwd.restype = None
        L.ksf_unitarize_bwd.argtypes = [_dp, _dp, _dp, C.c_double]
        L.ksl_hisq_links.restype = C.c_long
        L.ksl_hisq_links.argtypes = [_ip, _dp, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_double, C.c_double]

    @staticmethod
    def _dims(dims):
        return np.ascontiguousarray(dims, dtype=np.int32)

    def smear(self, dims, links, coeffs, want_long=True):
        links = np.ascontiguousarray(links, np.float64)
        fat = np.zeros_like(links)
        lng = np.zeros_like(links) if want_long else None
        self.lib.ksl_smear(self._dims(dims), np.ascontiguousarray(coeffs, np.float64), links, fat,
                           lng.ctypes.data if want_long else None)
        return fat, lng

    def unitarize(self, V, allow_svd=True, svd_rel=1e-8, svd_abs=1e-8):
        V = np.ascontiguousarray(V, np.float64)
        W = np.zeros_like(V)
        n = self.lib.ksl_unitarize(V, W, V.size // 18, int(allow_svd), svd_rel, svd_abs)
        return W, int(n)

    FORCE_FILTER = 5.0e-5   # HISQ_FORCE_FILTER of ks_imp_rhmc's build (ks_imp_rhmc/Make_template)

    def unitarize_bwd(self, V, GW, force_filter=FORCE_FILTER):
        """One link of the projection's reverse step: G_V from V and G_W, (3,3,2) arrays."""
        GV = np.zeros((3, 3, 2))
        self.lib.ksf_unitarize_bwd(np.ascontiguousarray(V, np.float64), np.ascontiguousarray(GW, np.float64), GV,
                                   force_filter)
        return GV

    def hisq_force(self, dims, links, multi_x, residues, eps, coeffs1=None, coeffs2=None, force_filter=FORCE_FILTER):
        """ks_force_oracle.c: the momentum update of eo_fermion_force_multi as (V,4,10) anti_hermitmat arrays.
        force_filter = 0 gives the unregularised derivative."""
        self.lib.ksf_set_force_filter(force_filter)
        links = np.ascontiguousarray(links, np.float64)
        xs = np.ascontiguousarray(multi_x, np.float64)
        res = np.ascontiguousarray(residues, np.float64)
        c1 = np.ascontiguousarray(self.FAT7 if coeffs1 is None else coeffs1, np.float64)
        c2 = np.ascontiguousarray(self.ASQTAD_LIKE if coeffs2 is None else coeffs2, np.float64)
        mom = np.zeros((links.shape[0], 4, 10))
        self.lib.ksf_hisq_force(self._dims(dims), c1, c2, links, xs, res, xs.shape[0], eps, mom)
        return mom

    NAIK_TABLE = (1.0 / 8.0, -1.0 / 24.0)   # one-link and Naik coefficients of the reference's third path table

    def hisq_force_naik(self, dims, links, multi_x, residues, n_orders, eps_naik, eps, coeffs1=None, coeffs2=None,
                        coeffs3=None, force_filter=FORCE_FILTER):
        """ksf_hisq_force_naik: several Naik epsilons; the terms come in len(n_orders) classes, class k solved with
        the links of eps_naik[k] (eps_naik[0] = 0)."""
        self.lib.ksf_set_force_filter(force_filter)
        links = np.ascontiguousarray(links, np.float64)
        xs = np.ascontiguousarray(multi_x, np.float64)
        assert xs.shape[0] == sum(n_orders) and len(n_orders) == len(eps_naik)
        c1 = np.ascontiguousarray(self.FAT7 if coeffs1 is None else coeffs1, np.float64)
        c2 = np.ascontiguousarray(self.ASQTAD_LIKE if coeffs2 is None else coeffs2, np.float64)
        c3 = np.ascontiguousarray(self.NAIK_TABLE if coeffs3 is None else coeffs3, np.float64)
        mom = np.zeros((links.shape[0], 4, 10))
        self.lib.ksf_hisq_force_naik(self._dims(dims), c1, c2, c3, links, xs, np.ascontiguousarray(residues, np.float64),
                                     len(n_orders), np.ascontiguousarray(n_orders, np.int32),
                                     np.ascontiguousarray(eps_naik, np.float64), eps, mom)
        return mom

    def hisq_links(self, dims, links, coeffs1=None, coeffs2=None, allow_svd=True, svd_rel=1e-8, svd_abs=1e-8):
        links = np.ascontiguousarray(links, np.float64)
        out = {k: np.zeros_like(links) for k in ("V", "W", "fat", "lng")}
        c1 = np.ascontiguousarray(self.FAT7 if coeffs1 is None else coeffs1, np.float64)
        c2 = np.ascontiguousarray(self.ASQTAD_LIKE if coeffs2 is None else coeffs2, np.float64)
        out["nsvd"] = int(self.lib.ksl_hisq_links(self._dims(dims), c1, c2, links, out["V"].ctypes.data,
                                                  out["W"].ctypes.data, out["fat"].ctypes.data,
                                                  out["lng"].ctypes.data, int(allow_svd), svd_rel, svd_abs))
        return out


def _qic(o):
    return dict(final_rsq=o[0], final_relrsq=o[1], size_r=o[2], size_relr=o[3],
                final_iters=int(o[4]), final_restart=int(o[5]), converged=int(o[6]))


def ref_path(variant=""):
    return os.path.join(HERE, "_ref", "libmilcref%s.so" % variant)


def ref_available(variant=""):
    return os.path.exists(ref_path(variant))


class MilcRef:
    """The reference's own compiled hot path (oracle/_ref).  One lattice geometry
    per process because MILC keeps its geometry in process globals."""

    def __init__(self, dims, variant=""):
        self.lib = C.CDLL(ref_path(variant))
        L = self.lib
        self.prec = L.milcref_precision()
        self.dtype = np.float64 if self.prec == 2 else np.float32
        rp = np.ctypeslib.ndpointer(dtype=self.dtype, flags="C_CONTIGUOUS")
        L.milcref_init.argtypes = [C.c_int] * 4
        L.milcref_sites_on_node.restype = C.c_long
        L.milcref_node_index.argtypes = [C.c_int] * 4
        L.milcref_set_links.argtypes = [rp, rp, C.c_double]
        L.milcref_dslash.restype = None
        L.milcref_dslash.argtypes = [rp, rp, C.c_int]
        L.milcref_congrad.argtypes = [rp, rp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_double, _dp]
        L.milcref_multicg.argtypes = [rp, rp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_double, _dp]
        L.milcref_time_dslash.restype = C.c_double
        L.milcref_time_dslash.argtypes = [rp, rp, C.c_int, C.c_int]
        ro = np.ctypeslib.ndpointer(dtype=self.dtype, flags="C_CONTIGUOUS")
        L.milcref_hisq_links.argtypes = [rp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _dp]
        L.milcref_smear.argtypes = [rp, _dp, ro, C.c_void_p]
        L.milcref_unitarize.argtypes = [rp, ro, C.c_long]
        L.milcref_mat_invert_uml.argtypes = [rp, ro, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, _dp]
        L.milcref_hisq_force.argtypes = [rp, rp, rp, C.c_int, C.c_double, ro]
        L.milcref_mat_invert_uml_deflated.argtypes = [rp, ro, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, rp, _dp, _dp]
        L.milcref_hisq_force_naik.argtypes = [rp, rp, rp, C.c_int, _ip, _dp, C.c_double, ro, C.c_void_p]
        self.has_eigcg = hasattr(L, "milcref_eigcg")   # inc_eigcg.c needs a LAPACK at build time (oracle/build_ref.sh)
        if self.has_eigcg:
            L.milcref_eigcg.argtypes = [rp, ro, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _dp, ro, _dp]
            L.milcref_inc_eigcg_init.argtypes = [C.c_int, C.c_int, C.c_int]
            L.milcref_inc_eigcg.argtypes = [rp, ro, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]
            L.milcref_eigcg_pairs.argtypes = [C.c_int, _dp, ro, C.c_void_p]
        self.has_meson = hasattr(L, "milcref_meson_cont_mom")
        if self.has_meson:
            L.milcref_set_ape_links.argtypes = [rp]
            L.milcref_spin_taste_index.argtypes = [C.c_char_p]
            L.milcref_spin_taste_op.restype = None
            L.milcref_spin_taste_op.argtypes = [C.c_int, _ip, ro, rp]
            L.milcref_meson_cont_mom.argtypes = [rp, rp, C.c_int, _ip, C.c_char_p, C.c_int, _ip, _ip, _ip, _dp, _ip, C.c_int, _ip, _dp]
        self.dims = tuple(int(d) for d in dims)
        if L.milcref_init(*self.dims) != 0:
            raise RuntimeError("MilcRef: process already initialised with another geometry")
        self.vol = L.milcref_sites_on_node()

    def node_index(self, x, y, z, t):
        return self.lib.milcref_node_index(x, y, z, t)

    def set_links(self, fat, lng, eps_naik=0.0):
        self.lib.milcref_set_links(np.ascontiguousarray(fat, self.dtype),
                                   np.ascontiguousarray(lng, self.dtype), eps_naik)

    def dslash(self, src, parity, dest=None):
        src = np.ascontiguousarray(src, self.dtype)
        if dest is None:
            dest = np.zeros_like(src)
        self.lib.milcref_dslash(src, dest, parity)
        return dest

    def congrad(self, src, dest, mass, parity, niter, nrestart, resid, relresid=0.0):
        out = np.zeros(7)
        it = self.lib.milcref_congrad(np.ascontiguousarray(src, self.dtype), dest, mass, parity,
                                      niter, nrestart, resid, relresid, out)
        return it, _qic(out)

    def multicg(self, src, offsets, parity, niter, nrestart, resid, relresid=0.0):
        offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        n = len(offsets)
        src = np.ascontiguousarray(src, self.dtype)
        psim = np.zeros((n,) + src.shape, dtype=self.dtype)
        out = np.zeros(7 * n)
        it = self.lib.milcref_multicg(src, psim, offsets, n, parity, niter, nrestart, resid,
                                      relresid, out)
        return it, psim, [_qic(out[7 * j:7 * j + 7]) for j in range(n)]

    # -- HISQ link construction --------------------------------
    def hisq_links(self, links):
        """The reference's create_hisq_links_milc on thin links (V,4,3,3,2) with KS phases in.
        Returns dict(V, W, fat, lng, coeffs[3][6], nsvd)."""
        links = np.ascontiguousarray(links, self.dtype)
        out = {k: np.zeros_like(links) for k in ("V", "W", "fat", "lng")}
        coeffs = np.zeros(18)
        nsvd = self.lib.milcref_hisq_links(links, out["V"].ctypes.data, out["W"].ctypes.data,
                                           out["fat"].ctypes.data, out["lng"].ctypes.data, coeffs)
        out["coeffs"] = coeffs.reshape(3, 6)
        out["nsvd"] = nsvd
        return out

    def smear(self, links, coeffs, want_long=True):
        """load_fatlinks_cpu (+ load_lnglinks) with {one_link, naik, 3-, 5-, 7-staple, lepage}."""
        links = np.ascontiguousarray(links, self.dtype)
        fat = np.zeros_like(links)
        lng = np.zeros_like(links) if want_long else None
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        self.lib.milcref_smear(links, c, fat, lng.ctypes.data if want_long else None)
        return fat, lng

    def unitarize(self, V):
        """u3_unitarize_analytic on every matrix of V (..., 3, 3, 2); returns (W, svd count)."""
        V = np.ascontiguousarray(V, self.dtype)
        W = np.zeros_like(V)
        n = self.lib.milcref_unitarize(V, W, V.size // 18)
        return W, n

    def mat_invert_uml(self, srcs, dsts, mass, niter, nrestart, resid):
        """mat_invert_uml_field (one source) / mat_invert_block_uml (several): srcs, dsts arrays of
        shape (nsrc, V, 3, 2); dsts = guesses in, solutions out.  Returns (iterations, qic dict)."""
        srcs = np.ascontiguousarray(srcs, self.dtype)
        out = np.zeros(7)
        it = self.lib.milcref_mat_invert_uml(srcs, dsts, srcs.shape[0], mass, niter, nrestart, resid, out)
        return it, _qic(out)

    def mat_invert_uml_deflated(self, src, dst, mass, niter, nrestart, resid, eigvec, eigval):
        """mat_invert_uml_field with qic->deflate = 1 and the given low modes (mat_invert.c:131-183,328-402)."""
        ev = np.ascontiguousarray(eigvec, self.dtype)
        out = np.zeros(7)
        it = self.lib.milcref_mat_invert_uml_deflated(np.ascontiguousarray(src, self.dtype), dst, mass, niter, nrestart, resid,
                                                      ev.shape[0], ev, np.ascontiguousarray(eigval, np.float64), out)
        return it, _qic(out)

    def hisq_force(self, links, multi_x, residues, eps):
        """eo_fermion_force_multi (generic_ks/fermion_force_hisq_multi.c:170-216) on thin links with
        phases in: returns the momentum update as (V,4,10) anti_hermitmat arrays
        {m01, m02, m12 (re,im), m00im, m11im, m22im, space} and the SVD/filter count."""
        links = np.ascontiguousarray(links, self.dtype)
        xs = np.ascontiguousarray(multi_x, self.dtype)
        res = np.ascontiguousarray(residues, self.dtype)
        mom = np.zeros((self.vol, 4, 10), dtype=self.dtype)
        n = self.lib.milcref_hisq_force(links, xs, res, xs.shape[0], eps, mom)
        return mom, n

    def hisq_force_naik(self, links, multi_x, residues, n_orders, eps_naik, eps, want_links=False):
        """The same with several Naik epsilons (fermion_force_hisq_multi.c:1285-1375): terms in classes of
        n_orders[k], class k belonging to eps_naik[k] (eps_naik[0] = 0).  want_links: also the (fat, long) links
        of every class, (n_naiks, 2, V, 4, 3, 3, 2)."""
        links = np.ascontiguousarray(links, self.dtype)
        xs = np.ascontiguousarray(multi_x, self.dtype)
        res = np.ascontiguousarray(residues, self.dtype)
        assert xs.shape[0] == sum(n_orders) and len(n_orders) == len(eps_naik)
        mom = np.zeros((self.vol, 4, 10), dtype=self.dtype)
        fl = np.zeros((len(n_orders), 2) + links.shape, dtype=self.dtype) if want_links else None
        n = self.lib.milcref_hisq_force_naik(links, xs, res, len(n_orders), np.ascontiguousarray(n_orders, np.int32),
                                             np.ascontiguousarray(eps_naik, np.float64), eps, mom,
                                             fl.ctypes.data if want_links else None)
        return (mom, n, fl) if want_links else (mom, n)

    # ---- meson tie-ups (generic_ks/ks_meson_mom.c, spin_taste_ops.c) ------------------------------------------
    def set_ape_links(self, links):
        """ks_spectrum's ape_links global: the links the one-link sink operators shift with."""
        self.lib.milcref_set_ape_links(np.ascontiguousarray(links, self.dtype))

    def spin_taste_index(self, label):
        return self.lib.milcref_spin_taste_index(label.encode())

    def spin_taste_op(self, index, r0, src):
        """spin_taste_op_fn on the links of set_links."""
        src = np.ascontiguousarray(src, self.dtype)
        dest = np.zeros_like(src)
        self.lib.milcref_spin_taste_op(int(index), np.ascontiguousarray(r0, dtype=np.int32), dest, src)
        return dest

    def meson_cont_mom(self, src1, src2, mom, mom_parity, spin_taste, p_index, phase, factor, corr_index, nprop, r0, prop=None):
        """ks_meson_cont_mom: prop[m][t] (complex), accumulated onto `prop` when given.  Correlators with the same sink
        operator must be consecutive (one group of the reference's corr_table each)."""
        mom = np.ascontiguousarray(mom, dtype=np.int32).reshape(-1, 3)
        par = np.ascontiguousarray(mom_parity, dtype=np.int8).reshape(-1, 3)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        out = np.zeros((nprop, self.dims[3], 2))
        if prop is not None:
            out[..., 0], out[..., 1] = prop.real, prop.imag
        self.lib.milcref_meson_cont_mom(np.ascontiguousarray(src1, self.dtype), np.ascontiguousarray(src2, self.dtype),
                                        mom.shape[0], mom, par.tobytes(), len(spin_taste), i32(spin_taste), i32(p_index),
                                        i32(phase), np.ascontiguousarray(factor, np.float64), i32(corr_index), nprop, i32(r0), out)
        return out[..., 0] + 1j * out[..., 1]

    # ---- eigCG (generic_ks/inc_eigcg.c) ---------------------------------------------------------------------
    def eigcg(self, src, dest, mass, parity, niter, nrestart, resid, m, nvecs):
        """ks_eigCG_parity: returns (iterations, eigVal[nvecs] of -D^2, eigVec (nvecs, V, 3, 2), qic)."""
        out, val = np.zeros(7), np.zeros(m)
        vec = np.zeros((max(nvecs, 1), self.vol, 3, 2))
        it = self.lib.milcref_eigcg(np.ascontiguousarray(src, self.dtype), dest, mass, parity, niter, nrestart, resid, m, nvecs,
                                    val, vec, out)
        return it, val[:nvecs].copy(), vec[:nvecs], _qic(out)

    def inc_eigcg_init(self, m, nvecs, nvecs_max):
        self._inc_max = nvecs_max
        self.lib.milcref_inc_eigcg_init(m, nvecs, nvecs_max)

    def inc_eigcg(self, src, dest, mass, parity, niter, nrestart, resid):
        """ks_inc_eigCG_parity: returns (iterations, qic, eigenvectors accumulated so far)."""
        out, n = np.zeros(7), C.c_int(0)
        it = self.lib.milcref_inc_eigcg(np.ascontiguousarray(src, self.dtype), dest, mass, parity, niter, nrestart, resid, out,
                                        C.byref(n))
        return it, _qic(out), n.value

    def eigcg_pairs(self, parity, ncurr):
        """calc_eigenpairs: (eigVal[ncurr] of -D^2 ascending, eigVec (ncurr, V, 3, 2))."""
        val = np.zeros(self._inc_max)
        vec = np.zeros((max(ncurr, 1), self.vol, 3, 2))
        n = self.lib.milcref_eigcg_pairs(parity, val, vec, None)
        return val[:n].copy(), vec[:n]

    def time_dslash(self, src, parity, ncalls):
        src = np.ascontiguousarray(src, self.dtype)
        dest = np.zeros_like(src)
        return self.lib.milcref_time_dslash(src, dest, parity, ncalls)

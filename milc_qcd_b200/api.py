"""Host-side (numpy) face of libb200ks, mirroring the reference's operator interface.

The names, argument meaning and outputs follow MILC's solver API for this path
(``include/imp_ferm_links.h:47-282``):

* :func:`dslash_fn_field`            <- ``dslash_fn_field(src, dest, parity, fn)``
* :func:`ks_congrad_parity_gpu`      <- ``ks_congrad_parity_gpu(src, dest, qic, mass, fn)``
* :func:`ks_congrad_field`           <- ``ks_congrad_field`` (EVEN / ODD / EVENANDODD fan-out,
  ``generic_ks/d_congrad5_fn.c:16-60``)
* :func:`ks_congrad_block_parity_gpu` <- ``ks_congrad_block_parity_gpu(nsrc, src[], dest[], qic, mass, fn)``
* :func:`ks_multicg_offset_field_gpu` <- ``ks_multicg_offset_field_gpu(src, psim, ksp, n, qic, fn)``

with :class:`quark_invert_control`, :class:`ks_param` and :class:`fn_links_t` mirroring
``include/generic_quark_types.h:131-139,167-190`` and ``include/fn_links.h:12-20``.  All of it
is a thin veneer over the C ABI in ``include/b200ks.h``; the compute is hand-written CUDA.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import EVEN, ODD, EVENANDODD, InvertArgs, InvertResult, check


@dataclass
class quark_invert_control:
    """include/generic_quark_types.h:167-190 (fields used by the KS solvers)."""
    prec: int = 2
    min: int = 0
    max: int = 500
    nrestart: int = 5
    parity: int = EVEN
    start_flag: int = 1
    nsrc: int = 1
    resid: float = 1e-10
    relresid: float = 0.0
    final_rsq: float = 0.0
    final_relrsq: float = 0.0
    size_r: float = 0.0
    size_relr: float = 0.0
    converged: int = 1
    final_iters: int = 0
    final_restart: int = 0
    # not in MILC: selects the inner precision like the HALF_MIXED / MAX_MIXED build macros
    mixed_precision: int = 0
    device_seconds: float = 0.0


@dataclass
class ks_param:
    """include/generic_quark_types.h:131-139."""
    mass: float = 0.0
    offset: float = 0.0
    residue: float = 0.0
    naik_term_epsilon: float = 0.0
    naik_term_epsilon_index: int = 0


@dataclass
class fn_links_t:
    """include/fn_links.h:12-20: fat[4*i+dir], lng[4*i+dir] as (V,4,3,3,2) arrays."""
    fat: np.ndarray = None
    lng: np.ndarray = None
    eps_naik: float = 0.0
    notify_quda_new_links: int = 1
    dims: tuple = field(default=None)


def _host_prec(a):
    if a.dtype == np.float64:
        return 2
    if a.dtype == np.float32:
        return 1
    raise TypeError("MILC fields are float32 (PRECISION=1) or float64 (PRECISION=2)")


def _ptr(a):
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous (MILC host layout)")
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One lattice on one GPU (``b200ks_create``)."""

    def __init__(self, dims, device=0, grid=None, rank=0, nranks=1, nccl_id=None, ngpu=None, devices=None):
        """Single GPU: Context(dims).  One rank per GPU: Context(global_dims, device, grid, rank,
        nranks, nccl_id) with nccl_id from :func:`comm_unique_id` on rank 0, broadcast by the
        caller; host arrays are then the LOCAL sub-lattice (see milc_qcd_b200.dist).
        One process, several GPUs: Context(dims, ngpu=N[, devices=[...]]) (``b200ks_create_multi``):
        host arrays stay MILC's GLOBAL arrays, exactly as for a single GPU."""
        self.lib = _lib.load()
        self.global_dims = tuple(int(d) for d in dims)
        arr = (C.c_int * 4)(*self.global_dims)
        if ngpu is not None and grid is None:
            self.dims = self.global_dims
            devs = None if devices is None else (C.c_int * int(ngpu))(*[int(d) for d in devices])
            self.h = self.lib.b200ks_create_multi(arr, int(ngpu), devs)
        elif grid is None:
            self.dims = self.global_dims
            self.h = self.lib.b200ks_create(arr, device)
        else:
            self.dims = tuple(self.global_dims[d] // int(grid[d]) for d in range(4))
            garr = (C.c_int * 4)(*[int(g) for g in grid])
            buf = C.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
            self.h = self.lib.b200ks_create_dist(arr, garr, rank, nranks, buf, device)
        if not self.h:
            raise _lib.B200KSError("b200ks_create failed: %s" % self.lib.b200ks_last_error().decode())
        self.volume = int(np.prod(self.dims))
        self._links_of = None

    def num_gpus(self):
        return self.lib.b200ks_num_gpus(self.h)

    def links_sync(self, fat, lng, changed_hint=0, mode=2):
        """b200ks_links_sync: what the MILC-facing shims call before every solve."""
        return check(self.lib.b200ks_links_sync(self.h, _ptr(fat), _ptr(lng), _host_prec(fat), int(changed_hint), int(mode)),
                     "b200ks_links_sync")

    def links_sync_stats(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        check(self.lib.b200ks_links_sync_stats(self.h, C.byref(a), C.byref(b)), "b200ks_links_sync_stats")
        return {"uploads": a.value, "verifications": b.value}

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200ks_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- links -----------------------------------------------------------------------------
    def long_link_info(self):
        """(complex numbers stored per long link: 9 or 7, worst load-time misfit)."""
        nc, dev = C.c_int(), C.c_double()
        check(self.lib.b200ks_long_link_info(self.h, C.byref(nc), C.byref(dev)), "b200ks_long_link_info")
        return nc.value, dev.value

    def load_links(self, fat, lng, long_recon=0):
        if fat.dtype != lng.dtype:
            raise TypeError("fat and lng must have the same precision")
        n = self.volume * 4 * 18
        if fat.size != n or lng.size != n:
            raise ValueError("links must hold su3_matrix[4*volume]")
        check(self.lib.b200ks_load_links(self.h, _ptr(fat), _ptr(lng), _host_prec(fat), long_recon),
              "b200ks_load_links")

    def ensure_links(self, fn):
        """Device link cache keyed like the QUDA seam (fn identity + notify flag,
        generic_ks/d_congrad5_fn_gpu.c:121-126)."""
        if self._links_of is not fn or fn.notify_quda_new_links:
            self.load_links(fn.fat, fn.lng)
            self._links_of = fn
            fn.notify_quda_new_links = 0

    # -- fermion-link construction -----------------------------------
    # {one_link, naik, three_staple, five_staple, seven_staple, lepage} of the reference's HISQ
    # action (generic_ks/imp_actions/hisq/hisq_u3_action.h:33-37,74-81)
    HISQ_FAT7 = (1.0 / 8.0, 0.0, -1.0 / 16.0, 1.0 / 64.0, -1.0 / 384.0, 0.0)
    HISQ_ASQTAD_LIKE = (1.0, -1.0 / 24.0, -1.0 / 16.0, 1.0 / 64.0, -1.0 / 384.0, -1.0 / 8.0)

    @staticmethod
    def _coeffs(c):
        if len(c) != 6:
            raise ValueError("path coefficients: {one_link, naik, three_staple, five_staple, seven_staple, lepage}")
        return (C.c_double * 6)(*[float(x) for x in c])

    def ks_links(self, links, coeffs, want_long=True):
        """load_fatlinks + load_lnglinks (qudaLoadKSLink): returns (fat, lng or None)."""
        fat = np.zeros_like(links)
        lng = np.zeros_like(links) if want_long else None
        check(self.lib.b200ks_ks_links(self.h, self._coeffs(coeffs), _ptr(links), _ptr(fat),
                                       _ptr(lng) if want_long else None, _host_prec(links)), "b200ks_ks_links")
        return fat, lng

    def unitarized_links(self, links, coeffs, want_v=True):
        """fat7 smear + U(3) projection (qudaLoadUnitarizedLink): returns (V or None, W, svd count)."""
        V = np.zeros_like(links) if want_v else None
        W = np.zeros_like(links)
        n = C.c_longlong(0)
        check(self.lib.b200ks_unitarized_links(self.h, self._coeffs(coeffs), _ptr(links), _ptr(V) if want_v else None,
                                               _ptr(W), _host_prec(links), C.byref(n)), "b200ks_unitarized_links")
        return V, W, n.value

    def hisq_links(self, links, coeffs1=None, coeffs2=None):
        """create_hisq_links_milc's chain U -> V -> W -> (fat, lng), intermediate fields resident."""
        out = {k: np.zeros_like(links) for k in ("V", "W", "fat", "lng")}
        n = C.c_longlong(0)
        check(self.lib.b200ks_hisq_links(self.h, self._coeffs(self.HISQ_FAT7 if coeffs1 is None else coeffs1),
                                         self._coeffs(self.HISQ_ASQTAD_LIKE if coeffs2 is None else coeffs2), _ptr(links), _ptr(out["V"]),
                                         _ptr(out["W"]), _ptr(out["fat"]), _ptr(out["lng"]), _host_prec(links),
                                         C.byref(n)), "b200ks_hisq_links")
        out["nsvd"] = n.value
        return out

    HISQ_FORCE_FILTER = 5.0e-5   # ks_imp_rhmc's build value (ks_imp_rhmc/Make_template)

    HISQ_NAIK_TABLE = (1.0 / 8.0, -1.0 / 24.0)   # one-link and Naik coefficients of the reference's third path table

    def hisq_force(self, U, V, W, multi_x, residues, eps, coeffs1=None, coeffs2=None, force_filter=HISQ_FORCE_FILTER,
                   n_orders=None, eps_naik=None, coeffs3=None):
        """b200ks_hisq_force with MILC's weights (one-hop 2 res_j, three-hop naik * 2 res_j,
        generic_ks/fermion_force_hisq_multi.c:2189-2192): returns the momentum increment as
        (V,4,10) anti_hermitmat arrays.  force_filter = 0: the unregularised derivative.
        Several Naik epsilons (:2196-2222): the terms come in len(n_orders) classes of n_orders[k] terms,
        class k solved with the links of eps_naik[k] (eps_naik[0] = 0)."""
        c1 = self.HISQ_FAT7 if coeffs1 is None else coeffs1
        c2 = self.HISQ_ASQTAD_LIKE if coeffs2 is None else coeffs2
        c3 = self.HISQ_NAIK_TABLE if coeffs3 is None else coeffs3
        n = len(multi_x)
        extra = []
        if n_orders is not None and len(n_orders) > 1:
            assert sum(n_orders) == n and len(eps_naik) == len(n_orders)
            j = n_orders[0]
            for k in range(1, len(n_orders)):
                for _ in range(n_orders[k]):
                    extra.append((float(c3[0]) * eps_naik[k] * 2.0 * float(residues[j]),
                                  float(c3[1]) * eps_naik[k] * 2.0 * float(residues[j])))
                    j += 1
        cf = (C.c_double * (2 * (n + len(extra))))()
        for j, r in enumerate(residues):
            cf[2 * j] = 2.0 * float(r)
            cf[2 * j + 1] = float(c2[1]) * 2.0 * float(r)
        for i, (a, b) in enumerate(extra):
            cf[2 * (n + i)] = a
            cf[2 * (n + i) + 1] = b
        ptrs = (C.c_void_p * n)(*[_ptr(x).value for x in multi_x])
        mom = np.zeros((self.volume, 4, 10), dtype=U.dtype)
        check(self.lib.b200ks_hisq_force(self.h, n, len(extra), cf, ptrs, self._coeffs(c2), self._coeffs(c1), _ptr(W), _ptr(V),
                                         _ptr(U), eps, float(force_filter), _ptr(mom), _host_prec(U)), "b200ks_hisq_force")
        return mom

    def hisq_links_time(self, seed, reps, coeffs1=None, coeffs2=None):
        ms, n = C.c_double(), C.c_longlong(0)
        check(self.lib.b200ks_hisq_links_time(self.h, self._coeffs(self.HISQ_FAT7 if coeffs1 is None else coeffs1),
                                              self._coeffs(self.HISQ_ASQTAD_LIKE if coeffs2 is None else coeffs2), seed, reps, C.byref(ms),
                                              C.byref(n)), "b200ks_hisq_links_time")
        return ms.value, n.value

    def hisq_links_fetch(self, which, dtype=np.float64):
        out = np.zeros((self.volume, 4, 3, 3, 2), dtype=dtype)
        check(self.lib.b200ks_hisq_links_fetch(self.h, which, _ptr(out), _host_prec(out)), "b200ks_hisq_links_fetch")
        return out

    # -- host-buffer operators ------------------------------------------------------------------
    def dslash(self, src, dest, parity):
        check(self.lib.b200ks_dslash(self.h, _ptr(src), _ptr(dest), parity, _host_prec(src)), "b200ks_dslash")
        return dest

    def congrad(self, src, dest, mass, parity, max_iter, nrestart, resid, relresid=0.0,
                mixed_precision=0, check_interval=0):
        args = InvertArgs(parity, max_iter, nrestart, resid, relresid, mixed_precision, check_interval)
        res = InvertResult()
        it = check(self.lib.b200ks_congrad(self.h, _ptr(src), _ptr(dest), mass, C.byref(args), C.byref(res),
                                           _host_prec(src)), "b200ks_congrad")
        return it, res.as_dict()

    def congrad_block(self, srcs, dests, mass, parity, max_iter, nrestart, resid, relresid=0.0,
                      mixed_precision=0, check_interval=0):
        """b200ks_congrad_block: len(srcs) systems, solved up to four at a time."""
        n = len(srcs)
        args = InvertArgs(parity, max_iter, nrestart, resid, relresid, mixed_precision, check_interval)
        res = (InvertResult * max(n, 1))()
        sp = (C.c_void_p * max(n, 1))(*[_ptr(a).value for a in srcs])
        dp = (C.c_void_p * max(n, 1))(*[_ptr(a).value for a in dests])
        prec = _host_prec(srcs[0]) if n else 2
        it = check(self.lib.b200ks_congrad_block(self.h, n, sp, dp, mass, C.byref(args), res, prec),
                   "b200ks_congrad_block")
        return it, [res[j].as_dict() for j in range(n)]

    def mat_invert_uml(self, srcs, dsts, mass, max_iter, nrestart, resid, relresid=0.0, mixed_precision=0):
        """b200ks_mat_invert_uml: dst_k = (D + 2m)^-1 src_k on both parities, resident sequence."""
        n = len(srcs)
        args = InvertArgs(EVEN, max_iter, nrestart, resid, relresid, mixed_precision, 0)
        res = (InvertResult * (2 * n))()
        sp = (C.c_void_p * n)(*[_ptr(a).value for a in srcs])
        dp = (C.c_void_p * n)(*[_ptr(a).value for a in dsts])
        it = check(self.lib.b200ks_mat_invert_uml(self.h, n, sp, dp, mass, C.byref(args), res, _host_prec(srcs[0])),
                   "b200ks_mat_invert_uml")
        return it, [(res[2 * k].as_dict(), res[2 * k + 1].as_dict()) for k in range(n)]

    def multicg_rational(self, src, offsets, parity, max_iter, nrestart, resid, residues=None, want_psim=True,
                         fill_other=False, mixed_precision=0):
        """b200ks_multicg_rational: returns (iterations, psim list or None, dest or None, results)."""
        n = len(offsets)
        args = InvertArgs(parity, max_iter, nrestart, resid, 0.0, mixed_precision, 0)
        res = (InvertResult * n)()
        offs = (C.c_double * n)(*[float(o) for o in offsets])
        psim = [np.zeros_like(src) for _ in range(n)] if want_psim else None
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in psim]) if want_psim else None
        dest = np.zeros_like(src) if residues is not None else None
        rr = (C.c_double * (n + 1))(*[float(r) for r in residues]) if residues is not None else None
        it = check(self.lib.b200ks_multicg_rational(self.h, _ptr(src), ptrs, _ptr(dest) if dest is not None else None,
                                                    offs, rr, n, int(fill_other), C.byref(args), res, _host_prec(src)),
                   "b200ks_multicg_rational")
        return it, psim, dest, [res[j].as_dict() for j in range(n)]

    def multicg(self, src, psim, offsets, parity, max_iter, nrestart, resid, mixed_precision=0,
                check_interval=0):
        n = len(offsets)
        args = InvertArgs(parity, max_iter, nrestart, resid, 0.0, mixed_precision, check_interval)
        res = (InvertResult * max(n, 1))()
        offs = (C.c_double * max(n, 1))(*[float(o) for o in offsets])
        ptrs = (C.c_void_p * max(n, 1))(*[p.ctypes.data for p in psim])
        it = check(self.lib.b200ks_multicg(self.h, _ptr(src), ptrs, offs, n, C.byref(args), res,
                                           _host_prec(src)), "b200ks_multicg")
        return it, [res[j].as_dict() for j in range(n)]

    # -- device-resident interface --------------------------------------------------------------
    def vec_create(self):
        return check(self.lib.b200ks_vec_create(self.h), "b200ks_vec_create")

    def vec_free(self, v):
        check(self.lib.b200ks_vec_free(self.h, v), "b200ks_vec_free")

    def vec_upload(self, v, host, parity=EVENANDODD):
        check(self.lib.b200ks_vec_upload(self.h, v, _ptr(host), parity, _host_prec(host)), "b200ks_vec_upload")

    def vec_download(self, v, host, parity=EVENANDODD):
        check(self.lib.b200ks_vec_download(self.h, v, _ptr(host), parity, _host_prec(host)), "b200ks_vec_download")
        return host

    def vec_zero(self, v, parity=EVENANDODD):
        check(self.lib.b200ks_vec_zero(self.h, v, parity), "b200ks_vec_zero")

    def vec_norm2(self, v, parity=EVENANDODD):
        out = C.c_double()
        check(self.lib.b200ks_vec_norm2(self.h, v, parity, C.byref(out)), "b200ks_vec_norm2")
        return out.value

    def eig_set(self, vecs, eigval, use_in_uml=True):
        """b200ks_eig_set: device vectors `vecs` (both parities uploaded) with eigenvalues `eigval` of -D_eo D_oe
        become the context's low-mode set (MILC's eigVec / eigVal); an empty list drops it."""
        n = len(vecs)
        hv = (C.c_int * max(n, 1))(*[int(v) for v in vecs])
        ev = (C.c_double * max(n, 1))(*[float(x) for x in eigval])
        check(self.lib.b200ks_eig_set(self.h, n, hv, ev, int(use_in_uml)), "b200ks_eig_set")

    def eig_count(self):
        return self.lib.b200ks_eig_count(self.h)

    # -- eigCG (generic_ks/inc_eigcg.c) ------------------------------------------------------------
    def eigcg_init(self, m, nvecs, nvecs_max):
        check(self.lib.b200ks_eigcg_init(self.h, m, nvecs, nvecs_max), "b200ks_eigcg_init")

    def inc_eigcg(self, src, dest, mass, parity, max_iter, nrestart, resid):
        """ks_inc_eigCG_parity: one solve of the incremental sequence; returns (iterations, result dict)."""
        args = InvertArgs(parity, max_iter, nrestart, resid, 0.0, 0, 0)
        res = InvertResult()
        it = check(self.lib.b200ks_inc_eigcg(self.h, _ptr(src), _ptr(dest), mass, C.byref(args), C.byref(res), _host_prec(src)),
                   "b200ks_inc_eigcg")
        return it, res.as_dict()

    def inc_eigcg_dev(self, vsrc, vdest, mass, parity, max_iter, nrestart, resid):
        args = InvertArgs(parity, max_iter, nrestart, resid, 0.0, 0, 0)
        res = InvertResult()
        it = check(self.lib.b200ks_inc_eigcg_dev(self.h, vsrc, vdest, mass, C.byref(args), C.byref(res)), "b200ks_inc_eigcg_dev")
        return it, res.as_dict()

    def eigcg_count(self):
        return self.lib.b200ks_eigcg_count(self.h)

    def eigcg_pairs(self):
        """calc_eigenpairs: Ritz values of -D^2 (ascending) of everything accumulated; rotates the vectors."""
        n = self.eigcg_count()
        out = (C.c_double * max(n, 1))()
        n = check(self.lib.b200ks_eigcg_pairs(self.h, out, n), "b200ks_eigcg_pairs")
        return np.array(out[:n])

    def eigcg_vec(self, j, dtype=np.float64):
        out = np.zeros((self.volume, 3, 2), dtype=dtype)
        check(self.lib.b200ks_eigcg_vec_download(self.h, j, _ptr(out), _host_prec(out)), "b200ks_eigcg_vec_download")
        return out

    # -- meson tie-ups (generic_ks/ks_meson_mom.c) --------------------------------------------------
    def _meson_args(self, r0, mom, mom_parity):
        mom = np.ascontiguousarray(mom, dtype=np.int32).reshape(-1, 3)
        par = np.ascontiguousarray(mom_parity, dtype=np.int8).reshape(-1, 3)
        assert par.shape == mom.shape
        nmom = mom.shape[0]
        r0a = (C.c_int * 4)(*[int(v) for v in r0])
        out = np.zeros((self.global_dims[3], nmom, 2))
        return nmom, r0a, mom, par, out

    def meson_mom(self, antiquark, quark, spin, r0, mom, mom_parity):
        """The site loops of ks_meson_cont_mom for one sink spin-taste assignment: corr[t][p] (complex), host fields in
        MILC's layout.  spin = gamma bits of a local sink operator (15 = pion5, 0 = pion05, ...) or -1 (already applied)."""
        nmom, r0a, mom, par, out = self._meson_args(r0, mom, mom_parity)
        assert _host_prec(antiquark) == _host_prec(quark)
        check(self.lib.b200ks_meson_mom(self.h, _ptr(antiquark), _ptr(quark), _host_prec(quark), int(spin), r0a, nmom,
                                        mom.ctypes.data_as(C.POINTER(C.c_int)), par.tobytes(),
                                        out.ctypes.data_as(C.POINTER(C.c_double))), "b200ks_meson_mom")
        return out[..., 0] + 1j * out[..., 1]

    def meson_mom_dev(self, vantiquark, vquark, spin, r0, mom, mom_parity):
        nmom, r0a, mom, par, out = self._meson_args(r0, mom, mom_parity)
        check(self.lib.b200ks_meson_mom_dev(self.h, vantiquark, vquark, int(spin), r0a, nmom,
                                            mom.ctypes.data_as(C.POINTER(C.c_int)), par.tobytes(),
                                            out.ctypes.data_as(C.POINTER(C.c_double))), "b200ks_meson_mom_dev")
        return out[..., 0] + 1j * out[..., 1]

    def deflate_dev(self, vsrc, vdst, mass, parity):
        """b200ks_deflate_dev: deflate() of generic_ks/mat_invert.c:131-183 on device vectors."""
        check(self.lib.b200ks_deflate_dev(self.h, vsrc, vdst, mass, parity), "b200ks_deflate_dev")

    def dslash_dev(self, vsrc, vdest, parity, prec=2):
        check(self.lib.b200ks_dslash_dev(self.h, vsrc, vdest, parity, prec), "b200ks_dslash_dev")

    def congrad_dev(self, vsrc, vdest, mass, parity, max_iter, nrestart, resid, relresid=0.0,
                    mixed_precision=0, check_interval=0):
        args = InvertArgs(parity, max_iter, nrestart, resid, relresid, mixed_precision, check_interval)
        res = InvertResult()
        it = check(self.lib.b200ks_congrad_dev(self.h, vsrc, vdest, mass, C.byref(args), C.byref(res)),
                   "b200ks_congrad_dev")
        return it, res.as_dict()

    def congrad_block_dev(self, vsrcs, vdests, mass, parity, max_iter, nrestart, resid, relresid=0.0,
                          mixed_precision=0, check_interval=0):
        n = len(vsrcs)
        args = InvertArgs(parity, max_iter, nrestart, resid, relresid, mixed_precision, check_interval)
        res = (InvertResult * max(n, 1))()
        vs = (C.c_int * max(n, 1))(*vsrcs)
        vd = (C.c_int * max(n, 1))(*vdests)
        it = check(self.lib.b200ks_congrad_block_dev(self.h, n, vs, vd, mass, C.byref(args), res),
                   "b200ks_congrad_block_dev")
        return it, [res[j].as_dict() for j in range(n)]

    def dslash_block_dev(self, vsrcs, vdests, parity, prec=2):
        n = len(vsrcs)
        vs = (C.c_int * n)(*vsrcs)
        vd = (C.c_int * n)(*vdests)
        check(self.lib.b200ks_dslash_block_dev(self.h, n, vs, vd, parity, prec), "b200ks_dslash_block_dev")

    def dslash_block_time(self, prec, nrhs, parity, n):
        out = C.c_double()
        check(self.lib.b200ks_dslash_block_time(self.h, prec, nrhs, parity, n, C.byref(out)), "b200ks_dslash_block_time")
        return out.value

    def multicg_dev(self, vsrc, vpsim, offsets, parity, max_iter, nrestart, resid, mixed_precision=0,
                    check_interval=0):
        n = len(offsets)
        args = InvertArgs(parity, max_iter, nrestart, resid, 0.0, mixed_precision, check_interval)
        res = (InvertResult * max(n, 1))()
        offs = (C.c_double * max(n, 1))(*[float(o) for o in offsets])
        hs = (C.c_int * max(n, 1))(*vpsim)
        it = check(self.lib.b200ks_multicg_dev(self.h, vsrc, hs, offs, n, C.byref(args), res), "b200ks_multicg_dev")
        return it, [res[j].as_dict() for j in range(n)]

    def vec_gaussian(self, v, parity, seed):
        check(self.lib.b200ks_vec_gaussian(self.h, v, parity, seed), "b200ks_vec_gaussian")

    def links_synthetic(self, seed, long_recon=0):
        check(self.lib.b200ks_links_synthetic(self.h, seed, long_recon), "b200ks_links_synthetic")

    def links_download(self, dtype=np.float64):
        fat = np.zeros((self.volume, 4, 3, 3, 2), dtype=dtype)
        lng = np.zeros_like(fat)
        check(self.lib.b200ks_links_download(self.h, _ptr(fat), _ptr(lng), _host_prec(fat)), "b200ks_links_download")
        return fat, lng

    def dslash_time(self, prec, parity, n):
        out = C.c_double()
        check(self.lib.b200ks_dslash_time(self.h, prec, parity, n, C.byref(out)), "b200ks_dslash_time")
        return out.value

    def halo_mode(self):
        """0 no partitioned direction, 1 NCCL send/recv halos, 2 peer-to-peer push halos."""
        return check(self.lib.b200ks_halo_mode(self.h), "b200ks_halo_mode")

    def launch_count(self):
        return int(self.lib.b200ks_launch_count(self.h))

    def device_bytes(self):
        return int(self.lib.b200ks_device_bytes(self.h))


def comm_unique_id():
    """128-byte ncclUniqueId for b200ks_create_dist (call on rank 0, broadcast to the others)."""
    buf = C.create_string_buffer(128)
    check(_lib.load().b200ks_comm_unique_id(buf), "b200ks_comm_unique_id")
    return bytes(buf.raw)


# ---- MILC-named operators ------------------------------------------------------------------------
_ctx_cache = {}


def _context_for(fn):
    dims = tuple(fn.dims)
    ctx = _ctx_cache.get(dims)
    if ctx is None:
        ctx = _ctx_cache[dims] = Context(dims)
    ctx.ensure_links(fn)
    return ctx


def finalize():
    """qudaFinalize analogue: drop cached contexts."""
    for ctx in _ctx_cache.values():
        ctx.close()
    _ctx_cache.clear()


def dslash_fn_field(src, dest, parity, fn):
    """generic_ks/dslash_fn.c:306-356: dest(parity) = D src."""
    if fn is None:
        raise ValueError("dslash_fn_field: invalid fn links!")
    ctx = _context_for(fn)
    if parity == EVENANDODD and src is dest:
        raise ValueError("in-place dslash needs a single parity")
    return ctx.dslash(src, dest, parity)


def _store(qic, res):
    for k in ("final_rsq", "final_relrsq", "size_r", "size_relr", "final_iters", "final_restart",
              "converged", "device_seconds"):
        setattr(qic, k, res[k])


def ks_congrad_parity_gpu(t_src, t_dest, qic, mass, fn):
    """generic_ks/d_congrad5_fn_gpu.c:35-172 with the CPU solver's semantics
    (generic_ks/d_congrad5_fn_milc.c:60-407).  Returns iterations, fills qic."""
    if fn is None:
        raise ValueError("ks_congrad_parity_gpu: Called with NULL fn")
    if qic.parity not in (EVEN, ODD):
        raise ValueError("ks_congrad_parity_gpu: Unrecognised parity")
    ctx = _context_for(fn)
    it, res = ctx.congrad(t_src, t_dest, mass, qic.parity, qic.max, qic.nrestart, qic.resid, qic.relresid,
                          qic.mixed_precision)
    _store(qic, res)
    return it


def ks_congrad_block_parity_gpu(nsrc, t_src, t_dest, qic, mass, fn):
    """generic_ks/d_congrad5_fn_gpu.c:175-312 (the CPU reference is a loop of single solves,
    generic_ks/d_congrad5_fn_milc.c:409-417).  Returns the total iterations; qic reports the
    worst right-hand side, as the QUDA glue does with the one residual it gets back."""
    if fn is None:
        raise ValueError("ks_congrad_block_parity_gpu: Called with NULL fn")
    if qic.parity not in (EVEN, ODD):
        raise ValueError("ks_congrad_block_parity_gpu: Unrecognised parity")
    ctx = _context_for(fn)
    it, res = ctx.congrad_block(t_src[:nsrc], t_dest[:nsrc], mass, qic.parity, qic.max, qic.nrestart, qic.resid,
                                qic.relresid, qic.mixed_precision)
    if res:
        worst = max(res, key=lambda r: r["final_rsq"])
        _store(qic, dict(worst, final_iters=it, converged=int(all(r["converged"] for r in res)),
                         final_restart=max(r["final_restart"] for r in res)))
    return it


def mat_invert_uml_field(src, dst, qic, mass, fn):
    """generic_ks/mat_invert.c:328-402 as one device-resident sequence (b200ks_mat_invert_uml):
    dst = (D + 2m)^-1 src on all sites.  Returns the iterations of both solves; qic->final_iters
    is their sum, the other outputs are the odd solve's, as in the reference."""
    if fn is None:
        raise ValueError("mat_invert_uml_field: Called with NULL fn")
    ctx = _context_for(fn)
    it, res = ctx.mat_invert_uml([src], [dst], mass, qic.max, qic.nrestart, qic.resid, qic.relresid, qic.mixed_precision)
    even, odd = res[0]
    _store(qic, dict(odd, final_iters=even["final_iters"] + odd["final_iters"]))
    qic.parity = ODD
    return it


def mat_invert_block_uml(src, dst, mass, nsrc, qic, fn):
    """generic_ks/mat_invert.c:409-475: the same sequence for nsrc sources through the block solver."""
    if fn is None:
        raise ValueError("mat_invert_block_uml: Called with NULL fn")
    ctx = _context_for(fn)
    it, res = ctx.mat_invert_uml(src[:nsrc], dst[:nsrc], mass, qic.max, qic.nrestart, qic.resid, qic.relresid,
                                 qic.mixed_precision)
    worst = max((r for pair in res for r in pair), key=lambda r: r["final_rsq"])
    _store(qic, dict(worst, final_iters=it, converged=int(all(r["converged"] for pair in res for r in pair))))
    qic.parity = ODD
    return it


def ks_congrad_field(src, dest, qic, mass, fn):
    """generic_ks/d_congrad5_fn.c:16-60: EVENANDODD = EVEN solve then ODD solve."""
    iters = 0
    want = qic.parity
    try:
        for par in ((EVEN, ODD) if want == EVENANDODD else (want,)):
            qic.parity = par
            iters += ks_congrad_parity_gpu(src, dest, qic, mass, fn)
    finally:
        qic.parity = want
    return iters


def ks_multicg_offset_field_gpu(src, psim, ksp, num_offsets, qic, fn):
    """generic_ks/ks_multicg_offset_gpu.c:38-252 with the CPU algorithm's semantics
    (generic_ks/ks_multicg_offset.c:63-505).  qic is a list with one entry per offset."""
    if num_offsets == 0:
        return 0
    if qic[0].relresid != 0.0:
        raise ValueError("ks_multicg_offset_field_gpu: GPU code does not yet support a Fermilab-type relative residual")
    if qic[0].parity == EVENANDODD:
        raise ValueError("ks_multicg_offset_field_gpu: EVENANDODD not supported")
    if fn is None:
        raise ValueError("ks_multicg_offset_field: Called with NULL fn")
    ctx = _context_for(fn)
    offsets = [ksp[j].offset for j in range(num_offsets)]
    it, res = ctx.multicg(src, psim[:num_offsets], offsets, qic[0].parity, qic[0].max, qic[0].nrestart,
                          qic[0].resid, qic[0].mixed_precision)
    for j in range(num_offsets):
        _store(qic[j], res[j])
    return it

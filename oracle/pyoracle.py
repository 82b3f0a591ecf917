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
    src = os.path.join(HERE, "ks_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
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
        L.kso_multicg.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, _dp]

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

    def multicg(self, dims, fat, lng, src, offsets, parity, niter, nrestart, resid, relresid=0.0):
        offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        n = len(offsets)
        psim = np.zeros((n,) + src.shape)
        out = np.zeros(7 * n)
        it = self.lib.kso_multicg(self._dims(dims), fat, lng, src, psim, offsets, n, parity, niter,
                                  nrestart, resid, relresid, out)
        return it, psim, [_qic(out[7 * j:7 * j + 7]) for j in range(n)]


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

    def time_dslash(self, src, parity, ncalls):
        src = np.ascontiguousarray(src, self.dtype)
        dest = np.zeros_like(src)
        return self.lib.milcref_time_dslash(src, dest, parity, ncalls)

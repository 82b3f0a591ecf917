"""ctypes binding of route 1, the COMPILED MILC-named symbols in libb200ks_milc.so
(include/b200ks_milc.h = MILC's own prototypes, include/imp_ferm_links.h:73-93,238-246):

    ks_congrad_parity_gpu, ks_congrad_block_parity_gpu, ks_multicg_offset_field_gpu,
    dslash_fn_field, mat_invert_uml_field_gpu, mat_invert_block_uml_gpu

This is what a MILC binary built with these objects executes; bench.py's end-to-end leg and
tests/test_gpu_seam.py call it exactly like MILC does: plain (pageable) host arrays in MILC's site order,
a quark_invert_control in, the same struct filled on return.  MILC_PRECISION = 2 layout.
"""
import ctypes as C
import os

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
EVEN, ODD, EVENANDODD = 2, 1, 3


class quark_invert_control(C.Structure):   # include/generic_quark_types.h:167-190, PRECISION 2
    _fields_ = [("prec", C.c_int), ("min", C.c_int), ("max", C.c_int), ("nrestart", C.c_int), ("parity", C.c_int),
                ("start_flag", C.c_int), ("nsrc", C.c_int), ("deflate", C.c_int), ("resid", C.c_double),
                ("relresid", C.c_double), ("mixed_rsq", C.c_double), ("final_rsq", C.c_double), ("final_relrsq", C.c_double),
                ("size_r", C.c_double), ("size_relr", C.c_double), ("converged", C.c_int), ("final_iters", C.c_int),
                ("final_restart", C.c_int), ("inv_type", C.c_int), ("mgparamfile", C.c_char * 256)]


class ks_param(C.Structure):               # include/generic_quark_types.h:131-139
    _fields_ = [("mass", C.c_double), ("charge", C.c_double), ("offset", C.c_double), ("residue", C.c_double),
                ("naik_term_epsilon_index", C.c_int), ("charge_index", C.c_int), ("naik_term_epsilon", C.c_double)]


class fn_links_t(C.Structure):             # include/fn_links.h:12-20
    _fields_ = [("phase", C.c_void_p), ("fat", C.c_void_p), ("lng", C.c_void_p), ("fatback", C.c_void_p),
                ("lngback", C.c_void_p), ("eps_naik", C.c_double), ("notify_quda_new_links", C.c_int)]


def qic(parity, resid, max_iter=500, nrestart=5, relresid=0.0, prec=2):
    q = quark_invert_control()
    q.prec, q.max, q.nrestart, q.parity, q.resid, q.relresid = prec, max_iter, nrestart, parity, resid, relresid
    return q


def fn_links(fat, lng, notify=1):
    fn = fn_links_t()
    fn.fat, fn.lng, fn.notify_quda_new_links = fat.ctypes.data, lng.ctypes.data, notify
    return fn


_shim = None


def load():
    """libb200ks_milc.so (MILC_PRECISION = 2) on top of the real libb200ks.so.  No fallback: a missing
    library raises, a missing GPU makes the first solver call terminate(1) like MILC's own glue."""
    global _shim
    if _shim is not None:
        return _shim
    _lib.load()
    path = os.path.join(HERE, "libb200ks_milc.so")
    if not os.path.exists(path):
        raise RuntimeError("milc_qcd_b200: %s is missing -- build it with `python -m milc_qcd_b200.build`" % path)
    lib = C.CDLL(path)
    Q, F = C.POINTER(quark_invert_control), C.POINTER(fn_links_t)
    lib.b200ks_milc_setup.argtypes = [C.c_int] * 5
    lib.b200ks_milc_context.restype = C.c_void_p
    lib.ks_congrad_parity_gpu.argtypes = [C.c_void_p, C.c_void_p, Q, C.c_double, F]
    lib.ks_congrad_block_parity_gpu.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), Q, C.c_double, F]
    lib.ks_multicg_offset_field_gpu.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(ks_param), C.c_int, Q, F]
    lib.dslash_fn_field.argtypes = [C.c_void_p, C.c_void_p, C.c_int, F]
    lib.mat_invert_uml_field_gpu.argtypes = [C.c_void_p, C.c_void_p, Q, C.c_double, F]
    lib.mat_invert_block_uml_gpu.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double, C.c_int, Q, F]
    _shim = lib
    return lib

"""ctypes binding of libb200ks.so (include/b200ks.h).  Fails loudly: if the shared
library is missing or no sm_100 GPU is usable there is no fallback of any kind."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200KS_LIB: load another build of the same library (kernel-variant probes under profiles/)
LIB_PATH = os.environ.get("B200KS_LIB") or os.path.join(HERE, "libb200ks.so")

EVEN, ODD, EVENANDODD = 2, 1, 3
PREC_HALF, PREC_SINGLE, PREC_DOUBLE = 0, 1, 2


class InvertArgs(C.Structure):
    _fields_ = [("parity", C.c_int), ("max_iter", C.c_int), ("nrestart", C.c_int),
                ("resid", C.c_double), ("relresid", C.c_double), ("mixed_precision", C.c_int),
                ("check_interval", C.c_int)]


class InvertResult(C.Structure):
    _fields_ = [("final_rsq", C.c_double), ("final_relrsq", C.c_double), ("size_r", C.c_double),
                ("size_relr", C.c_double), ("final_iters", C.c_int), ("final_restart", C.c_int),
                ("converged", C.c_int), ("device_seconds", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/b200ks.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("b200ks_version", C.c_int, []),
    ("b200ks_last_error", C.c_char_p, []),
    ("b200ks_device_count", C.c_int, []),
    ("b200ks_create", C.c_void_p, [C.POINTER(C.c_int), C.c_int]),
    ("b200ks_create_dist", C.c_void_p, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                        C.c_void_p, C.c_int]),
    ("b200ks_comm_unique_id", C.c_int, [C.c_void_p]),
    ("b200ks_create_multi", C.c_void_p, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    ("b200ks_num_gpus", C.c_int, [C.c_void_p]),
    ("b200ks_links_sync", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("b200ks_links_sync_stats", C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    ("b200ks_destroy", None, [C.c_void_p]),
    ("b200ks_load_links", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("b200ks_fingerprint", C.c_ulonglong, [C.c_void_p, C.c_size_t]),
    ("b200ks_long_link_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    ("b200ks_dslash", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("b200ks_congrad", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                 C.POINTER(InvertArgs), C.POINTER(InvertResult), C.c_int]),
    ("b200ks_multicg", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double),
                                 C.c_int, C.POINTER(InvertArgs), C.POINTER(InvertResult), C.c_int]),
    ("b200ks_vec_create", C.c_int, [C.c_void_p]),
    ("b200ks_vec_free", C.c_int, [C.c_void_p, C.c_int]),
    ("b200ks_vec_upload", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    ("b200ks_vec_download", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    ("b200ks_vec_zero", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("b200ks_vec_gaussian", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_ulonglong]),
    ("b200ks_vec_norm2", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("b200ks_links_synthetic", C.c_int, [C.c_void_p, C.c_ulonglong, C.c_int]),
    ("b200ks_links_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    ("b200ks_dslash_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("b200ks_congrad_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(InvertArgs),
                                     C.POINTER(InvertResult)]),
    ("b200ks_multicg_dev", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                     C.c_int, C.POINTER(InvertArgs), C.POINTER(InvertResult)]),
    ("b200ks_congrad_block", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double,
                                       C.POINTER(InvertArgs), C.POINTER(InvertResult), C.c_int]),
    ("b200ks_congrad_block_dev", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_double,
                                           C.POINTER(InvertArgs), C.POINTER(InvertResult)]),
    ("b200ks_dslash_block_dev", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]),
    ("b200ks_dslash_block_time", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("b200ks_mat_invert_uml", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double,
                                        C.POINTER(InvertArgs), C.POINTER(InvertResult), C.c_int]),
    ("b200ks_mat_invert_uml_dev", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_double,
                                            C.POINTER(InvertArgs), C.POINTER(InvertResult)]),
    ("b200ks_multicg_rational", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(InvertArgs),
                                          C.POINTER(InvertResult), C.c_int]),
    ("b200ks_hisq_force", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_void_p), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int]),
    ("b200ks_eig_set", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int]),
    ("b200ks_eig_count", C.c_int, [C.c_void_p]),
    ("b200ks_eig_use_in_uml", C.c_int, [C.c_void_p, C.c_int]),
    ("b200ks_deflate_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    ("b200ks_eigcg_init", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("b200ks_inc_eigcg", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(InvertArgs), C.POINTER(InvertResult), C.c_int]),
    ("b200ks_inc_eigcg_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(InvertArgs), C.POINTER(InvertResult)]),
    ("b200ks_eigcg_pairs", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    ("b200ks_eigcg_count", C.c_int, [C.c_void_p]),
    ("b200ks_eigcg_vec_download", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("b200ks_eigcg_hmatrix", C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    ("b200ks_meson_mom_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int),
                                       C.c_char_p, C.POINTER(C.c_double)]),
    ("b200ks_meson_mom", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                   C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_double)]),
    ("b200ks_ks_links", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    ("b200ks_unitarized_links", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_longlong)]),
    ("b200ks_hisq_links", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]),
    ("b200ks_hisq_links_time", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_ulonglong, C.c_int,
                                         C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    ("b200ks_hisq_links_fetch", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("b200ks_dslash_time", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("b200ks_halo_mode", C.c_int, [C.c_void_p]),
    ("b200ks_call_profile", C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    ("b200ks_launch_count", C.c_longlong, [C.c_void_p]),
    ("b200ks_stream", C.c_void_p, [C.c_void_p]),
    ("b200ks_device_bytes", C.c_size_t, [C.c_void_p]),
]

_lib = None


def load():
    """Load libb200ks.so and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "milc_qcd_b200: %s is missing -- build it with `python -m milc_qcd_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class B200KSError(RuntimeError):
    pass


def check(rc, what=""):
    if rc < 0:
        msg = load().b200ks_last_error().decode("utf-8", "replace")
        raise B200KSError("%s failed (%d): %s" % (what or "libb200ks call", rc, msg))
    return rc

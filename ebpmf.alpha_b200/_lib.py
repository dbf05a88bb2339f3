"""ctypes loader for lib/libebpmf_b200.so -- the plain C ABI of include/ebpmf_b200.h.

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is
usable, every solver call raises (``EbpmfError``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EBPMF_B200_LIB") or os.path.join(_HERE, "lib", "libebpmf_b200.so")   # override: A/B builds

EBPMF_OK = 0
ERR_NAMES = {1: "EBPMF_ERR_ARG", 2: "EBPMF_ERR_CUDA", 3: "EBPMF_ERR_NAN", 4: "EBPMF_ERR_VAR_TYPE",
             5: "EBPMF_ERR_STATE", 6: "EBPMF_ERR_NCCL", 7: "EBPMF_ERR_NOMEM"}
VAR_TYPE = {"constant": 0, "by_row": 1, "by_col": 2}        # R/ebpmf_log.R:178-192
OP_SCALAR, OP_ROWVEC, OP_COLVEC, OP_MATRIX = 0, 1, 2, 3


class EbpmfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class SolveInfo(C.Structure):
    _fields_ = [("n_updates", C.c_int32), ("converged", C.c_int32), ("passes", C.c_int32),
                ("reserved", C.c_int32), ("sym", C.c_double), ("ssexp", C.c_double),
                ("slogv", C.c_double), ("kernel_ms", C.c_double), ("pass1_ms", C.c_double),
                ("pass2_ms", C.c_double), ("reduce_ms", C.c_double)]


_vp, _i64, _i32, _dbl, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int
_PI = C.POINTER(SolveInfo)

# name -> (restype, argtypes): one row per declaration in include/ebpmf_b200.h
SIGNATURES = {
    "ebpmf_version": (C.c_char_p, []),
    "ebpmf_host_alloc": (_vp, [C.c_size_t]),
    "ebpmf_host_free": (_int, [_vp]),
    "ebpmf_host_pool_trim": (None, []),
    "ebpmf_host_pool_stats": (None, [_vp]),
    "ebpmf_device_count": (_int, [C.POINTER(_int)]),
    "ebpmf_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "ebpmf_ctx_destroy": (_int, [_vp]),
    "ebpmf_last_error": (C.c_char_p, [_vp]),
    "ebpmf_nccl_unique_id": (_int, [_vp]),
    "ebpmf_ctx_comm_init": (_int, [_vp, _int, _int, _vp]),
    "ebpmf_ctx_set_y_csc": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "ebpmf_ctx_set_y_dense": (_int, [_vp, _i64, _i64, _vp]),
    "ebpmf_ctx_sum_lfactorial": (_int, [_vp, C.POINTER(_dbl)]),
    "ebpmf_sum_lfactorial": (_int, [_vp, _i64, _vp, C.POINTER(_dbl)]),
    "ebpmf_ctx_set_m": (_int, [_vp, _vp]),
    "ebpmf_ctx_set_factors": (_int, [_vp, _i64, _vp, _vp]),
    "ebpmf_ctx_set_sigma2": (_int, [_vp, _int, _vp]),
    "ebpmf_ctx_vga_step": (_int, [_vp, _int, _dbl, _int, _PI]),
    "ebpmf_ctx_get_m": (_int, [_vp, _vp]),
    "ebpmf_ctx_peer_count": (_int, [_vp, C.POINTER(_int)]),
    "ebpmf_ctx_rewind_m": (_int, [_vp]),
    "ebpmf_ctx_get_v": (_int, [_vp, _vp]),
    "ebpmf_ctx_get_v_sum": (_int, [_vp, _vp]),
    "ebpmf_ctx_get_shape": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "ebpmf_ctx_get_y_csc": (_int, [_vp, _vp, _vp, _vp]),
    "ebpmf_ctx_get_factors": (_int, [_vp, _vp, _vp]),
    "ebpmf_ctx_get_sigma2": (_int, [_vp, _vp]),
    "ebpmf_ctx_get_rows": (_int, [_vp, _int, _i64, _i64, _vp]),
    "ebpmf_ctx_get_block": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _vp]),
    "ebpmf_ctx_update_sigma2": (_int, [_vp, _vp, _dbl, _dbl, _dbl, _i64, _i64, _vp]),
    "ebpmf_update_sigma2": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _dbl, _dbl, _dbl, _i64, _vp, _vp, _vp]),
    "ebpmf_calc_obj": (_int, [_vp, _i64, _i64, _dbl, _dbl, _dbl, _i64, _vp, _i64, _vp, _i64, _vp, _dbl, _dbl, _dbl, _dbl,
                              _dbl, C.POINTER(_dbl)]),
    "ebpmf_vga_step_host": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _int, _vp, _int, _dbl, _vp, _vp, _vp, _PI]),
    "ebpmf_vga_step_resident_host": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _int, _vp, _int, _dbl, _vp, _vp, _vp, _PI]),
    "ebpmf_ctx_m_products": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ebpmf_ctx_counters": (_int, [_vp, _vp, _int]),
    "ebpmf_group_create": (_int, [_int, _vp, C.POINTER(_vp)]),
    "ebpmf_group_destroy": (_int, [_vp]),
    "ebpmf_group_last_error": (C.c_char_p, [_vp]),
    "ebpmf_group_size": (_int, [_vp]),
    "ebpmf_group_set_option": (_int, [_vp, C.c_char_p, _dbl]),
    "ebpmf_group_set_y_csc": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "ebpmf_group_sum_lfactorial": (_int, [_vp, C.POINTER(_dbl)]),
    "ebpmf_group_set_m": (_int, [_vp, _i64, _i64, _vp]),
    "ebpmf_group_vga_step_host": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _int, _vp, _int, _dbl, _vp, _vp, _vp, _PI]),
    "ebpmf_batched_create": (_int, [_int, _int, C.POINTER(_vp)]),
    "ebpmf_batched_create_multi": (_int, [_int, _vp, _int, C.POINTER(_vp)]),
    "ebpmf_batched_destroy": (_int, [_vp]),
    "ebpmf_batched_last_error": (C.c_char_p, [_vp]),
    "ebpmf_batched_set_option": (_int, [_vp, C.c_char_p, _dbl]),
    "ebpmf_batched_set_y_csc": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "ebpmf_batched_sum_lfactorial": (_int, [_vp, C.POINTER(_dbl)]),
    "ebpmf_batched_vga_step_host": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _int, _vp, _int, _dbl, _vp, _vp, _vp, _PI, _vp,
                                           _vp]),
    "ebpmf_calc_split_obj": (_int, [_vp, _i64, _i64, _vp, _vp, _int, _vp, _vp, _int, _i64, _vp, _i64, _vp, _dbl, _dbl,
                                    C.POINTER(_dbl)]),
    "ebpmf_vga_newton": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _int, _vp, _int, _int,
                                _dbl, _vp, _vp, _PI]),
    "ebpmf_vga_fixed_iter": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _int, _vp, _int, _vp, _int, _int, _vp, _vp]),
    "ebpmf_vga_low_memory": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int,
                                    _int, _dbl, _vp, _PI]),
    "ebpmf_ctx_sim_data_log": (_int, [_vp, _i64, _i64, _i64, _i64, C.c_uint64, _dbl, _dbl, _dbl, _dbl, _dbl, _int,
                                      C.POINTER(_i64)]),
    "ebpmf_ctx_set_option": (_int, [_vp, C.c_char_p, _dbl]),
    "ebpmf_ctx_last_diag": (_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ebpmf_probe_fp64_tflops": (_int, [_vp, C.POINTER(_dbl)]),
}

_LIB = None


def load():
    """dlopen the C-ABI library and attach the prototypes.  Raises if it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EbpmfError(2, "CUDA extension %s is not built (run `make -C ebpmf.alpha_b200/csrc` or "
                                "__graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(rc, ctx=None):
    if rc != EBPMF_OK:
        msg = load().ebpmf_last_error(ctx)
        raise EbpmfError(rc, msg.decode() if msg else "")

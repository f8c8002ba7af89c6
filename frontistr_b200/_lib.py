"""ctypes loader for the product library frontistr_b200/libfistr_b200.so (C-ABI: include/fistr_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry fails
loudly, and every compute call fails with FISTR_B200_ERR_CUDA when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libfistr_b200.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)

_lib = None


class FistrB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fistr_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise FistrB200Error(-2, "%s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                     "there is no CPU fallback" % SO_PATH)
        _lib = declare(C.CDLL(SO_PATH))
    return _lib


def declare(L):
    """ctypes prototypes of the entry points whose defaults (int arguments, int result) do not fit"""
    L.fistr_b200_last_error.restype = C.c_char_p
    L.fistr_b200_launch_count.restype = C.c_int64
    L.fistr_b200_get_option.restype = C.c_int64
    L.fistr_b200_set_option.argtypes = [C.c_char_p, C.c_int64]
    L.fistr_b200_get_option.argtypes = [C.c_char_p]
    L.fistr_b200_host_register.argtypes = [C.c_void_p, C.c_int64]
    L.fistr_b200_host_unregister.argtypes = [C.c_void_p]
    L.fistr_b200_precond_33_setup.argtypes = [C.c_int, C.c_double, C.c_int]
    L.fistr_b200_mat_dump.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.fistr_b200_debug_format_e20.argtypes = [C.c_double, C.c_char_p]
    return L


def check(rc: int, allow_solver_codes: bool = False) -> int:
    if rc < 0 or (rc > 0 and not allow_solver_codes):
        raise FistrB200Error(rc, lib().fistr_b200_last_error().decode(errors="replace"))
    return rc


def ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None and a.size else C.cast(None, c_int_p)


def dp(a):
    return a.ctypes.data_as(c_dbl_p) if a is not None and a.size else C.cast(None, c_dbl_p)


EXPORTED = [
    "fistr_b200_init", "fistr_b200_finalize", "fistr_b200_last_error", "fistr_b200_set_option",
    "fistr_b200_get_option", "fistr_b200_comm_get_unique_id", "fistr_b200_comm_init",
    "fistr_b200_comm_finalize", "fistr_b200_mat_con_count", "fistr_b200_mat_con_fill",
    "fistr_b200_mat_structure", "fistr_b200_mat_upload", "fistr_b200_mat_download", "fistr_b200_mat_dump", "fistr_b200_debug_format_e20",
    "fistr_b200_host_register", "fistr_b200_host_unregister",
    "fistr_b200_stiffmatrix_3d", "fistr_b200_mat_ass_bc", "fistr_b200_nl_state_reset",
    "fistr_b200_nl_state_download", "fistr_b200_nl_stiffmatrix_3d", "fistr_b200_nl_update_3d", "fistr_b200_update_3d",
    "fistr_b200_nodal_stress_3d", "fistr_b200_solve_33",
    "fistr_b200_solve_history", "fistr_b200_solve_resid2", "fistr_b200_solve_profile", "fistr_b200_matvec_33", "fistr_b200_matresid_33",
    "fistr_b200_matvec_bench", "fistr_b200_precond_33_setup", "fistr_b200_precond_33_apply",
    "fistr_b200_precond_33_clear", "fistr_b200_precond_33_info", "fistr_b200_update_3_R",
    "fistr_b200_launch_count", "fistr_b200_device_info", "fistr_b200_fp64_peak",
    "fistr_b200_hbm_copy_peak", "fistr_b200_part_rcb", "fistr_b200_part_metis", "fistr_b200_part_metis_available", "fistr_b200_part_local_count", "fistr_b200_part_local_fill", "fistr_b200_debug_ssor_order", "fistr_b200_debug_sell_layout",
]

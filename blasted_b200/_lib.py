"""ctypes loader for libblasted_b200.so (the C ABI declared in include/blasted_b200.h).

There is no CPU fallback: if the shared library is missing, importing this module raises, and
every compute entry point fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB: development override (A/B of two builds inside one GPU call)
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libblasted_b200.so")


class Settings(C.Structure):
    """b200_settings (mirrors AsyncSolverSettings, include/solverfactory.hpp:46-68)."""
    _fields_ = [("prectype", C.c_int), ("bs", C.c_int), ("blockstorage", C.c_int),
                ("relax", C.c_int), ("thread_chunk_size", C.c_int), ("scale", C.c_int),
                ("nbuildsweeps", C.c_int), ("napplysweeps", C.c_int), ("fact_inittype", C.c_int),
                ("apply_inittype", C.c_int), ("compute_precinfo", C.c_int), ("level_mode", C.c_int)]


class SolveInfo(C.Structure):
    """b200_solve_info (mirrors SolveInfo, tests/solvers.hpp:19-27)."""
    _fields_ = [("converged", C.c_int), ("iters", C.c_int), ("resnorm", C.c_double),
                ("bnorm", C.c_double), ("device_ms", C.c_double), ("prec_ms", C.c_double)]


# name -> (restype, argtypes); every symbol include/blasted_b200.h declares
_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (_i, []),
    "b200_set_device": (_i, [_i]),
    "b200_kernel_launches": (_ll, []),
    "b200_reset_kernel_launches": (None, []),
    "b200_profile_enable": (None, [_i]),
    "b200_profile_reset": (None, []),
    "b200_profile_get": (_i, [_vp, _vp]),
    "b200_profile_classes": (_i, []),
    "b200_profile_get_n": (_i, [_i, _vp, _vp]),
    "b200_mat_create_host": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _pp]),
    "b200_mat_create_device": (_i, [_i, _i, _i, _vp, _vp, _vp, _pp]),
    "b200_mat_update_values_host": (_i, [_vp, _vp]),
    "b200_mat_update_values_device": (_i, [_vp, _vp]),
    "b200_mat_destroy": (None, [_vp]),
    "b200_mat_dim": (_i, [_vp]),
    "b200_mat_nbrows": (_i, [_vp]),
    "b200_mat_nnzb": (_ll, [_vp]),
    "b200_mat_set_stream": (_i, [_vp, _vp]),
    "b200_mat_release_workspace": (_i, [_vp]),
    "b200_mat_create_coo": (_i, [_i, _ll, _vp, _vp, _vp, _i, _i, _i, _pp]),
    "b200_mat_get_host": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "b200_mat_reorder": (_i, [_vp, _vp, _vp, _i, _i]),
    "b200_mat_scale": (_i, [_vp, _vp, _vp, _i, _i]),
    "b200_vec_reorder": (_i, [_vp, _ll, _i, _vp, _i, _i]),
    "b200_vec_scale": (_i, [_vp, _ll, _i, _vp, _i, _i]),
    "b200_mat_apply": (_i, [_vp, _vp, _vp]),
    "b200_mat_apply_host": (_i, [_vp, _vp, _vp]),
    "b200_mat_gemv3": (_i, [_vp, _d, _vp, _d, _vp, _vp]),
    "b200_mat_gemv3_host": (_i, [_vp, _d, _vp, _d, _vp, _vp]),
    "b200_prec_create": (_i, [C.POINTER(Settings), _vp, _pp]),
    "b200_prec_compute": (_i, [_vp, _vp]),
    "b200_prec_compute_host": (_i, [_vp, _vp, _vp]),
    "b200_prec_apply": (_i, [_vp, _vp, _vp]),
    "b200_prec_apply_host": (_i, [_vp, _vp, _vp]),
    "b200_prec_set_apply_params": (_i, [_vp, _d, _d, _d, _i, _i]),
    "b200_prec_apply_relax": (_i, [_vp, _vp, _vp, _i]),
    "b200_prec_apply_relax_host": (_i, [_vp, _vp, _vp, _i]),
    "b200_prec_check": (_i, [_vp]),
    "b200_prec_dim": (_i, [_vp]),
    "b200_prec_relaxation_available": (_i, [_vp]),
    "b200_prec_destroy": (None, [_vp]),
    "b200_prec_set_stream": (_i, [_vp, _vp]),
    "b200_prec_set_sweeps": (_i, [_vp, _i, _i]),
    "b200_prec_positions_size": (_i, [_vp, C.POINTER(_ll)]),
    "b200_prec_get_positions": (_i, [_vp, _vp, _vp, _vp]),
    "b200_prec_pattern_stats": (_i, [_vp, _vp]),
    "b200_prec_levels_size": (_i, [_vp, C.POINTER(_i)]),
    "b200_prec_get_levels": (_i, [_vp, _vp, _vp]),
    "b200_prec_get_factor": (_i, [_vp, _vp]),
    "b200_prec_get_dblocks": (_i, [_vp, _vp]),
    "b200_prec_get_scale": (_i, [_vp, _vp]),
    "b200_prec_ilu_residual": (_i, [_vp, C.POINTER(_d)]),
    "b200_prec_last_times": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "b200_solve": (_i, [C.c_char_p, _vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_solve_host": (_i, [C.c_char_p, _vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(SolveInfo)]),
    "b200_nccl_load": (_i, [C.c_char_p]),
    "b200_comm_unique_id": (_i, [_vp]),
    "b200_comm_create": (_i, [_vp, _i, _i, _pp]),
    "b200_comm_destroy": (None, [_vp]),
    "b200_comm_allreduce_sum": (_i, [_vp, _vp, _i]),
    "b200_dist_mat_create": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _pp]),
    "b200_dist_mat_destroy": (None, [_vp]),
    "b200_dist_mat_apply": (_i, [_vp, _vp, _vp]),
    "b200_dist_mat_apply_with_halo": (_i, [_vp, _vp, _vp, _vp]),
    "b200_dist_solve": (_i, [C.c_char_p, _vp, _vp, _vp, _vp, _d, _i, _i, C.POINTER(SolveInfo)]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C blasted_b200/csrc).  blasted_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _f = getattr(lib, _name)          # AttributeError here = header/library mismatch: fail loudly
    _f.restype = _res
    _f.argtypes = _args


def last_error() -> str:
    return lib.b200_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(last_error())

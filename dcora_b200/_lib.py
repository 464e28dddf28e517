"""ctypes binding of libdcora_b200.so (the C-ABI in include/dcora_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcora_b200.so")

c_int32_p = C.POINTER(C.c_int32)
c_double_p = C.POINTER(C.c_double)
c_uint8_p = C.POINTER(C.c_uint8)


class AgentDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("l", C.c_int32), ("b", C.c_int32),
                ("rowptr", c_int32_p), ("col", c_int32_p), ("val", c_double_p),
                ("G", c_double_p), ("precond_reg", C.c_double)]


class ROptParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("verbose", C.c_int32),
                ("gradnorm_tol", C.c_double), ("RGD_stepsize", C.c_double),
                ("RGD_use_preconditioner", C.c_int32), ("RTR_iterations", C.c_int32),
                ("RTR_tCG_iterations", C.c_int32), ("RTR_initial_radius", C.c_double),
                ("time_bound_s", C.c_double), ("retraction", C.c_int32),
                ("stable_decrease", C.c_int32)]


class ROptResult(C.Structure):
    _fields_ = [("success", C.c_int32), ("tCGStatus", C.c_int32),
                ("fInit", C.c_double), ("gradNormInit", C.c_double),
                ("fOpt", C.c_double), ("gradNormOpt", C.c_double),
                ("elapsedMs", C.c_double), ("outer_iterations", C.c_int32),
                ("inner_iterations", C.c_int32), ("accepted", C.c_int32),
                ("reserved", C.c_int32)]


# numpy view of an array of ROptResult
import numpy as _np
ROPT_RESULT_DTYPE = _np.dtype([("success", "<i4"), ("tCGStatus", "<i4"), ("fInit", "<f8"),
                               ("gradNormInit", "<f8"), ("fOpt", "<f8"), ("gradNormOpt", "<f8"),
                               ("elapsedMs", "<f8"), ("outer_iterations", "<i4"),
                               ("inner_iterations", "<i4"), ("accepted", "<i4"),
                               ("reserved", "<i4")])
assert ROPT_RESULT_DTYPE.itemsize == C.sizeof(ROptResult)


class HaloSegment(C.Structure):
    _fields_ = [("n", C.c_int32), ("l", C.c_int32), ("b", C.c_int32)]


class CertifyResult(C.Structure):
    _fields_ = [("is_psd", C.c_int32), ("converged", C.c_int32),
                ("matvecs", C.c_int32), ("restarts", C.c_int32),
                ("lambda_min", C.c_double), ("theta", C.c_double),
                ("residual", C.c_double)]


# every symbol include/dcora_b200.h declares: name -> (restype, argtypes)
VP = C.c_void_p
SYMBOLS = {
    "dcora_ropt_params_default": (None, [C.POINTER(ROptParams)]),
    "dcora_last_error": (C.c_char_p, []),
    "dcora_device_count": (C.c_int, []),
    "dcora_problem_create": (C.c_int, [C.POINTER(AgentDesc), C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(VP)]),
    "dcora_problem_destroy": (None, [VP]),
    "dcora_problem_clone_rank": (C.c_int, [VP, C.c_int32, C.POINTER(VP)]),
    "dcora_problem_dim": (C.c_int64, [VP]),
    "dcora_problem_rank": (C.c_int32, [VP]),
    "dcora_problem_num_agents": (C.c_int32, [VP]),
    "dcora_problem_nnz": (C.c_int64, [VP]),
    "dcora_problem_factor_nnz": (C.c_int64, [VP]),
    "dcora_problem_precond_reg": (C.c_double, [VP, C.c_int32]),
    "dcora_problem_stream": (VP, [VP]),
    "dcora_problem_launch_count": (C.c_int64, [VP]),
    "dcora_problem_set_G": (C.c_int, [VP, VP]),
    "dcora_problem_f": (C.c_int, [VP, VP, c_double_p]),
    "dcora_problem_egrad": (C.c_int, [VP, VP, VP]),
    "dcora_problem_hessvec": (C.c_int, [VP, VP, VP]),
    "dcora_problem_precond": (C.c_int, [VP, VP, VP, VP]),
    "dcora_problem_rgrad": (C.c_int, [VP, VP, VP]),
    "dcora_problem_rgradnorm": (C.c_int, [VP, VP, c_double_p]),
    "dcora_problem_retract": (C.c_int, [VP, VP, VP, VP, C.c_int32]),
    "dcora_problem_project_tangent": (C.c_int, [VP, VP, VP, VP]),
    "dcora_problem_project_manifold": (C.c_int, [VP, VP, VP]),
    "dcora_rtr_solve": (C.c_int, [VP, VP, VP, C.POINTER(ROptParams), c_uint8_p,
                                  C.POINTER(ROptResult)]),
    "dcora_certify": (C.c_int, [VP, VP, C.c_double, C.c_int32, C.c_double,
                                C.POINTER(CertifyResult), VP]),
    "dcora_problem_escape_saddle": (C.c_int, [VP, VP, C.c_double, VP, C.c_double,
                                              C.c_double, C.c_int32, C.c_int32, VP]),
    "dcora_problem_set_coupling": (C.c_int, [VP, C.POINTER(HaloSegment), C.c_int32,
                                             c_int32_p, c_int32_p, c_double_p]),
    "dcora_problem_set_coupling_keyed": (C.c_int, [VP, C.POINTER(HaloSegment), C.c_int32,
                                                   c_int32_p, c_int32_p, c_double_p,
                                                   C.POINTER(C.c_int64)]),
    "dcora_problem_halo_dim": (C.c_int64, [VP]),
    "dcora_problem_update_G": (C.c_int, [VP, VP, VP, c_uint8_p]),
    "dcora_problem_pack_columns": (C.c_int, [VP, VP, c_int32_p, C.c_int64, VP]),
    "dcora_problem_set_iterate_tracking": (C.c_int, [VP, C.c_int32]),
    "dcora_problem_load_iterate": (C.c_int, [VP, VP, VP]),
    "dcora_problem_store_iterate": (C.c_int, [VP, VP, VP]),
    "dcora_project_to_rotation_group": (C.c_int, [c_double_p, C.c_int32, c_double_p]),
    "dcora_problem_round_solution": (C.c_int, [VP, VP, VP]),
    "dcora_problem_align_trajectory": (C.c_int, [VP, VP, c_double_p, C.c_int32, VP]),
    "dcora_problem_certificate_prepare": (C.c_int, [VP, VP, VP]),
    "dcora_problem_certificate_apply": (C.c_int, [VP, VP, VP, VP]),
    "dcora_problem_cost_terms": (C.c_int, [VP, VP, c_double_p, c_double_p]),
    "dcora_problem_combine_project": (C.c_int, [VP, C.c_double, VP, C.c_double, VP,
                                                c_uint8_p, VP]),
    "dcora_problem_combine": (C.c_int, [VP, C.c_double, VP, C.c_double, VP, VP]),
    "dcora_problem_time_spmm": (C.c_int, [VP, C.c_int32, c_double_p, c_double_p]),
    "dcora_problem_time_kernel": (C.c_int, [VP, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "dcora_debug_trsv_trace": (C.c_int, [VP, C.c_int32, c_uint8_p, C.POINTER(C.c_uint64), C.c_int32]),
    "dcora_set_trsv_max_warps": (None, [C.c_int32]),
    "dcora_selfcheck_cholesky_threads": (C.c_int, [C.c_int32, c_int32_p, c_int32_p, c_double_p,
                                                   C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                                   c_double_p, C.POINTER(C.c_int64)]),
    "dcora_selfcheck_coupling": (C.c_int, [C.POINTER(AgentDesc), C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.POINTER(HaloSegment), C.c_int32,
                                           c_int32_p, c_int32_p, c_double_p,
                                           C.POINTER(C.c_int64), c_double_p,
                                           C.POINTER(C.c_int64)]),
    "dcora_selfcheck_layout": (C.c_int, [C.POINTER(AgentDesc), C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, c_double_p]),
    "dcora_selfcheck_cholesky": (C.c_int, [C.c_int32, c_int32_p, c_int32_p, c_double_p,
                                           C.c_int32, C.c_int32, C.c_double, c_double_p,
                                           C.POINTER(C.c_int64)]),
    "dcora_selfcheck_trsv": (C.c_int, [C.c_int32, c_int32_p, c_int32_p, c_double_p,
                                       C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                       c_double_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32)]),
}

_lib = None


class DcoraError(RuntimeError):
    pass


def load():
    """Load the shared library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DcoraError(
            f"{LIB_PATH} is missing: build it with `python -m dcora_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise DcoraError(f"dcora_b200 error {rc}: {load().dcora_last_error().decode()}")
    return rc

"""ctypes binding of libsypha_b200.so (include/sypha_b200.h).

The library is the product; there is NO Python/CPU fallback.  Importing this module without the
built shared object raises, and creating a workspace without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SYPHA_B200_LIB", _HERE / "lib" / "libsypha_b200.so"))

SB200_OK, SB200_ERR_INVALID, SB200_ERR_CUDA, SB200_ERR_NOMEM, SB200_ERR_NUMERICAL, SB200_ERR_UNSUPPORTED = range(6)
STRATEGY_AUTO, STRATEGY_CHOLESKY, STRATEGY_SYRK, STRATEGY_PCG = range(4)
STRATEGY_NAMES = {"auto": 0, "cholesky": 1, "syrk": 2, "pcg": 3, "krylov": 3, "dense": 1}
TRACE_COLS = 8


class sb200_caps(C.Structure):
    _fields_ = [("m_max", C.c_int), ("n_max", C.c_int), ("nnz_max", C.c_longlong)]


class sb200_params(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int), ("eta", C.c_double), ("mu_tol", C.c_double),
        ("gap_enabled", C.c_int), ("gap_window", C.c_int), ("gap_min_improv_pct", C.c_double),
        ("strategy", C.c_int),
        ("cg_max_iter", C.c_int), ("cg_tol_initial", C.c_double), ("cg_tol_final", C.c_double),
        ("cg_tol_decay", C.c_double),
        ("stop_flag", C.POINTER(C.c_int)), ("poll_every", C.c_int), ("use_graph", C.c_int),
        ("stop_cb", C.c_void_p), ("stop_user", C.c_void_p),
    ]


class sb200_result(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("reason", C.c_int), ("iterations", C.c_int),
        ("primal_obj", C.c_double), ("dual_obj", C.c_double), ("rel_gap", C.c_double), ("mu", C.c_double),
        ("ms_start", C.c_double), ("ms_setup", C.c_double), ("ms_loop", C.c_double),
        ("strategy_used", C.c_int), ("cg_iterations", C.c_longlong), ("kernels_launched", C.c_longlong),
        ("x_host", C.c_void_p), ("y_host", C.c_void_p), ("s_host", C.c_void_p),
        ("x0_host", C.c_void_p), ("y0_host", C.c_void_p), ("s0_host", C.c_void_p),
        ("xys_device", C.c_void_p),
    ]


class sb200_node_delta(C.Structure):
    _fields_ = [("n_extra_rows", C.c_int), ("var", C.POINTER(C.c_int)), ("coef", C.POINTER(C.c_double)),
                ("rhs", C.POINTER(C.c_double)),
                ("warm_start", C.c_void_p), ("warm_n", C.c_int), ("warm_m", C.c_int), ("warm_floor", C.c_double),
                ("export_xys", C.c_void_p)]


class sb200_heur_result(C.Structure):
    _fields_ = [("feasible", C.c_int), ("n_chosen", C.c_int), ("branch_var", C.c_int), ("repair_steps", C.c_int),
                ("cover_obj", C.c_double), ("branch_frac", C.c_double), ("rounded_obj", C.c_double),
                ("nif_feasible", C.c_int), ("reserved", C.c_int), ("nif_obj", C.c_double)]


class sb200_scp_model(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("n_orig", C.c_int), ("nnz", C.c_longlong),
                ("csr_offs", C.POINTER(C.c_int)), ("csr_inds", C.POINTER(C.c_int)), ("csr_vals", C.POINTER(C.c_double)),
                ("c", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double))]


class sb200_row_model(C.Structure):
    _fields_ = [("n_vars", C.c_int), ("n_rows", C.c_int), ("row_offs", C.POINTER(C.c_int)), ("row_inds", C.POINTER(C.c_int)),
                ("row_vals", C.POINTER(C.c_double)), ("row_lb", C.POINTER(C.c_double)), ("row_ub", C.POINTER(C.c_double)),
                ("obj", C.POINTER(C.c_double)), ("maximize", C.c_int)]


NEXT_NODE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(sb200_node_delta))
NODE_DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(sb200_result), C.POINTER(sb200_heur_result))


# every symbol include/sypha_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
SYMBOLS = {
    "sb200_version": (_i, []),
    "sb200_device_count": (_i, []),
    "sb200_ws_create": (_i, [_i, C.POINTER(sb200_caps), C.POINTER(_vp)]),
    "sb200_ws_destroy": (_i, [_vp]),
    "sb200_last_error": (C.c_char_p, [_vp]),
    "sb200_default_params": (None, [C.POINTER(sb200_params)]),
    "sb200_read_scp": (_i, [C.c_char_p, C.POINTER(sb200_scp_model)]),
    "sb200_free_scp": (None, [C.POINTER(sb200_scp_model)]),
    "sb200_standard_form_size": (_i, [C.POINTER(sb200_row_model), C.POINTER(_i), C.POINTER(_i), C.POINTER(_ll)]),
    "sb200_build_standard_form": (_i, [C.POINTER(sb200_row_model), C.POINTER(_i), C.POINTER(_i), C.POINTER(_d), C.POINTER(_d),
                                       C.POINTER(_d)]),
    "sb200_load_model": (_i, [_vp, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "sb200_solve": (_i, [_vp, C.POINTER(sb200_params), C.POINTER(sb200_result)]),
    "sb200_set_node_delta": (_i, [_vp, C.POINTER(sb200_node_delta)]),
    "sb200_solve_batch": (_i, [C.POINTER(_vp), _i, C.POINTER(sb200_node_delta), C.POINTER(sb200_params),
                               C.POINTER(sb200_result)]),
    "sb200_node_heuristics": (_i, [C.POINTER(_vp), _i, C.POINTER(sb200_heur_result)]),
    "sb200_set_heuristic_rules": (_i, [_vp, _i, _i, C.c_double]),
    "sb200_get_rounded": (_i, [_vp, _vp]),
    "sb200_get_cover": (_i, [_vp, _vp]),
    "sb200_set_concurrency_hint": (_i, [_vp, _i]),
    "sb200_set_solver_form": (_i, [_vp, _i]),
    "sb200_last_window": (_i, [_vp, C.POINTER(_d), C.POINTER(_i)]),
    "sb200_prepare_nodes": (_i, [_vp, _i]),
    "sb200_window_begin": (_i, [C.POINTER(_vp), _i, C.POINTER(sb200_node_delta), C.POINTER(sb200_params), C.POINTER(sb200_result), _i]),
    "sb200_window_finish": (_i, [C.POINTER(_vp), _i, C.POINTER(sb200_result), C.POINTER(sb200_heur_result)]),
    "sb200_solve_stream": (_i, [C.POINTER(_vp), _i, C.POINTER(sb200_params), NEXT_NODE_FN, NODE_DONE_FN, _vp]),
    "sb200_get_trace": (_i, [_vp, _vp, _i]),
    "sb200_get_device_iterates": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "sb200_get_iterates": (_i, [_vp, _vp, _vp, _vp]),
    "sb200_model_info": (_i, [_vp, C.POINTER(_ll), _i]),
    "sb200_stream": (_vp, [_vp]),
    "sb200_time_phase": (_i, [_vp, _i, _i, C.POINTER(_d)]),
    "sb200_ws_spmv": (_i, [_vp, _i, _vp, _vp]),
    "sb200_k_elem_min_mult": (_i, [_vp, _vp, _vp, _i, _vp]),
    "sb200_k_corrector_rhs": (_i, [_vp, _vp, _d, _d, _vp, _i, _vp]),
    "sb200_k_alpha_max": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "sb200_k_spmv_csr": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _d, _d, _vp]),
    "sb200_k_spmv_csc": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _d, _d, _vp]),
    "sb200_k_jacobi_diag": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb200_k_potrf": (_i, [_i, _vp, _i, _vp, _vp]),
    "sb200_k_potrs": (_i, [_i, _vp, _i, _vp, _vp]),
    "sb200_k_syrk": (_i, [_i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "sb200_assemble_normal": (_i, [_vp, _vp, _vp, _i]),
}

_lib = None


def load():
    """dlopen the CUDA library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with ./build.sh (or __graft_entry__.build()). "
            "sypha_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

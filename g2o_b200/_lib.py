"""ctypes loader for libg2ocu.so (the C ABI in include/g2ocu.h).  Fails loudly when the library is missing:
there is no CPU or PyTorch fallback behind this package."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libg2ocu.so")
_LIB = None

OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NUMERIC, E_COMM = 0, -1, -2, -3, -4, -5
ALGORITHM_GN, ALGORITHM_LM, ALGORITHM_DOGLEG = 0, 1, 2
LINEAR_PCG, LINEAR_DENSE = 0, 1
RESULT_OK, RESULT_TERMINATE, RESULT_FAIL = 1, 2, -1


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("linear_solver", ctypes.c_int32), ("pcg_tolerance", ctypes.c_double),
                ("pcg_max_iterations", ctypes.c_int32), ("pcg_absolute_tolerance", ctypes.c_int32), ("stream", ctypes.c_void_p)]


class IterationStats(ctypes.Structure):
    _fields_ = [("iteration", ctypes.c_int32), ("result", ctypes.c_int32), ("levenberg_iterations", ctypes.c_int32),
                ("iterations_linear_solver", ctypes.c_int32), ("chi2", ctypes.c_double), ("lambda_", ctypes.c_double),
                ("time_residuals", ctypes.c_double), ("time_quadratic_form", ctypes.c_double),
                ("time_schur_complement", ctypes.c_double), ("time_linear_solver", ctypes.c_double),
                ("time_linear_solution", ctypes.c_double), ("time_update", ctypes.c_double), ("time_iteration", ctypes.c_double),
                ("hessian_pose_dimension", ctypes.c_int64), ("hessian_landmark_dimension", ctypes.c_int64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["lambda"] = d.pop("lambda_")
        return d


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p)

# every symbol include/g2ocu.h declares
EXPORTS = ["g2ocu_default_config", "g2ocu_version", "g2ocu_last_error", "g2ocu_create", "g2ocu_destroy", "g2ocu_set_graph",
           "g2ocu_set_property", "g2ocu_set_force_stop_flag", "g2ocu_set_shard", "g2ocu_nccl_unique_id", "g2ocu_set_shard_nccl", "g2ocu_p2p_export", "g2ocu_p2p_import", "g2ocu_p2p_export_schur", "g2ocu_p2p_import_schur", "g2ocu_initialize_optimization", "g2ocu_init", "g2ocu_build_structure",
           "g2ocu_compute_active_errors", "g2ocu_active_robust_chi2", "g2ocu_active_chi2", "g2ocu_build_system",
           "g2ocu_set_lambda", "g2ocu_restore_diagonal", "g2ocu_solve", "g2ocu_update", "g2ocu_push", "g2ocu_pop",
           "g2ocu_discard_top", "g2ocu_compute_lambda_init", "g2ocu_compute_scale", "g2ocu_multiply_hessian", "g2ocu_compute_marginals",
           "g2ocu_solver_iteration", "g2ocu_optimize", "g2ocu_vector_size", "g2ocu_set_estimates", "g2ocu_get_estimates", "g2ocu_set_estimates_owned", "g2ocu_get_estimates_owned",
           "g2ocu_get_i32", "g2ocu_get_f64", "g2ocu_launch_count", "g2ocu_phase_time", "g2ocu_reset_counters",
           "g2ocu_linear_create", "g2ocu_linear_destroy", "g2ocu_linear_last_error", "g2ocu_linear_init", "g2ocu_linear_set_property", "g2ocu_linear_solve"]


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(g2o_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    P = ctypes.POINTER
    sig = {
        "g2ocu_default_config": (None, [P(Config)]), "g2ocu_version": (ctypes.c_int, []),
        "g2ocu_last_error": (ctypes.c_char_p, [vp]), "g2ocu_create": (ctypes.c_int, [P(Config), P(vp)]),
        "g2ocu_destroy": (None, [vp]), "g2ocu_set_graph": (ctypes.c_int, [vp, vp]),
        "g2ocu_set_property": (ctypes.c_int, [vp, ctypes.c_char_p, dbl]),
        "g2ocu_set_force_stop_flag": (ctypes.c_int, [vp, vp]),
        "g2ocu_set_shard": (ctypes.c_int, [vp, i32, i32, ALLREDUCE_FN, vp]),
        "g2ocu_nccl_unique_id": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p]),
        "g2ocu_set_shard_nccl": (ctypes.c_int, [vp, i32, i32, ctypes.c_char_p, ctypes.c_char_p]),
        "g2ocu_p2p_export": (ctypes.c_int, [vp, ctypes.c_char_p]), "g2ocu_p2p_import": (ctypes.c_int, [vp, ctypes.c_char_p]),
        "g2ocu_p2p_export_schur": (ctypes.c_int, [vp, ctypes.c_char_p]), "g2ocu_p2p_import_schur": (ctypes.c_int, [vp, ctypes.c_char_p]),
        "g2ocu_initialize_optimization": (ctypes.c_int, [vp, i32]), "g2ocu_init": (ctypes.c_int, [vp, i32]),
        "g2ocu_build_structure": (ctypes.c_int, [vp]), "g2ocu_compute_active_errors": (ctypes.c_int, [vp]),
        "g2ocu_active_robust_chi2": (ctypes.c_int, [vp, P(dbl)]), "g2ocu_active_chi2": (ctypes.c_int, [vp, P(dbl)]),
        "g2ocu_build_system": (ctypes.c_int, [vp]), "g2ocu_set_lambda": (ctypes.c_int, [vp, dbl, i32]),
        "g2ocu_restore_diagonal": (ctypes.c_int, [vp]), "g2ocu_solve": (ctypes.c_int, [vp, P(i32)]),
        "g2ocu_update": (ctypes.c_int, [vp, vp]), "g2ocu_push": (ctypes.c_int, [vp]), "g2ocu_pop": (ctypes.c_int, [vp]),
        "g2ocu_discard_top": (ctypes.c_int, [vp]), "g2ocu_compute_lambda_init": (ctypes.c_int, [vp, P(dbl)]),
        "g2ocu_compute_scale": (ctypes.c_int, [vp, dbl, P(dbl)]), "g2ocu_multiply_hessian": (ctypes.c_int, [vp, vp, vp]),
        "g2ocu_compute_marginals": (ctypes.c_int, [vp, i32, vp, vp, vp, P(i32)]),
        "g2ocu_solver_iteration": (ctypes.c_int, [vp, i32, i32, P(IterationStats)]),
        "g2ocu_optimize": (ctypes.c_int, [vp, i32, i32, P(IterationStats), P(i32)]),
        "g2ocu_vector_size": (i64, [vp]), "g2ocu_set_estimates": (ctypes.c_int, [vp, vp]),
        "g2ocu_get_estimates": (ctypes.c_int, [vp, vp]), "g2ocu_set_estimates_owned": (ctypes.c_int, [vp, vp]), "g2ocu_get_estimates_owned": (ctypes.c_int, [vp, vp]), "g2ocu_get_i32": (i64, [vp, ctypes.c_char_p, vp, i64]),
        "g2ocu_get_f64": (i64, [vp, ctypes.c_char_p, vp, i64]), "g2ocu_launch_count": (i64, [vp]),
        "g2ocu_phase_time": (ctypes.c_int, [vp, ctypes.c_char_p, P(dbl), P(i64), P(i64)]), "g2ocu_reset_counters": (ctypes.c_int, [vp]),
        "g2ocu_linear_create": (ctypes.c_int, [P(Config), P(vp)]), "g2ocu_linear_destroy": (None, [vp]), "g2ocu_linear_last_error": (ctypes.c_char_p, [vp]),
        "g2ocu_linear_init": (ctypes.c_int, [vp]), "g2ocu_linear_set_property": (ctypes.c_int, [vp, ctypes.c_char_p, dbl]),
        "g2ocu_linear_solve": (ctypes.c_int, [vp, i32, i32, vp, vp, vp, vp, vp, P(i32), P(i32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)          # AttributeError here = header/library mismatch
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L

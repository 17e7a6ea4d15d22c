"""ctypes view of the C-ABI in include/lvslam_b200.h.  Loads the in-tree liblvslam_b200.so and nothing else:
there is no Python or CPU implementation behind these calls, a missing library is a hard error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblvslam_b200.so")

LVS_KDTREE, LVS_DIRECT26, LVS_DIRECT7, LVS_DIRECT1 = 0, 1, 2, 3
LVS_NDT_OMP, LVS_NDT_PCA, LVS_NDT_GROUND = 0, 1, 2
LVS_ACC_EXACT, LVS_ACC_FAST = 0, 1

STATUS = {0: "LVS_OK", -1: "LVS_ERR_INVALID_ARG", -2: "LVS_ERR_NO_DEVICE", -3: "LVS_ERR_CUDA", -4: "LVS_ERR_OOM",
          -5: "LVS_ERR_NO_TARGET", -6: "LVS_ERR_NO_SOURCE", -7: "LVS_ERR_GRID_OVERFLOW", -8: "LVS_ERR_BAD_SLOT",
          -9: "LVS_ERR_NOT_SPD", -10: "LVS_ERR_EMPTY_GRAPH", -11: "LVS_ERR_PEER"}


class LvsError(RuntimeError):
    def __init__(self, status, detail):
        super().__init__("%s (%d): %s" % (STATUS.get(status, "?"), status, detail))
        self.status = status


class NdtParams(ctypes.Structure):
    _fields_ = [("resolution", ctypes.c_float), ("step_size", ctypes.c_double), ("outlier_ratio", ctypes.c_double),
                ("transformation_epsilon", ctypes.c_double), ("max_iterations", ctypes.c_int32), ("search_method", ctypes.c_int32),
                ("variant", ctypes.c_int32), ("min_points_per_voxel", ctypes.c_int32), ("min_covar_eigvalue_mult", ctypes.c_double),
                ("accumulation", ctypes.c_int32), ("lean_final_evaluation", ctypes.c_int32)]


class NdtResult(ctypes.Structure):
    _fields_ = [("final_transformation", ctypes.c_float * 16), ("converged", ctypes.c_int32), ("iterations", ctypes.c_int32),
                ("trans_probability", ctypes.c_double), ("n_eval", ctypes.c_int32), ("n_hess", ctypes.c_int32), ("score", ctypes.c_double)]


class NdtTraceRec(ctypes.Structure):
    _fields_ = [("p_before", ctypes.c_double * 6), ("dir", ctypes.c_double * 6), ("step", ctypes.c_double), ("score", ctypes.c_double),
                ("p_after", ctypes.c_double * 6), ("trials", ctypes.c_int32), ("hessian_recomputed", ctypes.c_int32)]


class PgoStats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("status", ctypes.c_int32), ("chi2_before", ctypes.c_double), ("chi2_after", ctypes.c_double),
                ("robust_chi2_after", ctypes.c_double), ("lambda_final", ctypes.c_double), ("device_ms", ctypes.c_double),
                ("linearize_ms", ctypes.c_double), ("solve_ms", ctypes.c_double), ("lm_trials", ctypes.c_int32), ("pcg_iterations", ctypes.c_int32),
                ("launches", ctypes.c_int32), ("linearize_launches", ctypes.c_int32)]


class PgoIterRec(ctypes.Structure):
    _fields_ = [("chi2", ctypes.c_double), ("lam", ctypes.c_double), ("trials", ctypes.c_int32), ("pcg_iterations", ctypes.c_int32)]

LVS_PGO_LM_CHOL, LVS_PGO_GN_CHOL, LVS_PGO_LM_PCG, LVS_PGO_GN_PCG = 0, 1, 2, 3


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("liblvslam_b200.so is not built (run `python -m lv_slam_b200.build`); there is no fallback path")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.lvs_last_error.restype = ctypes.c_char_p
        L.lvs_status_string.restype = ctypes.c_char_p
        L.lvs_status_string.argtypes = [i32]
        L.lvs_device_count.restype = i32
        L.lvs_ndt_default_params.restype = None
        L.lvs_ndt_default_params.argtypes = [vp]
        for name, args in {
            "lvs_ndt_create": [vp, i32, vp, vp], "lvs_ndt_destroy": [vp], "lvs_ndt_set_params": [vp, vp], "lvs_ndt_get_params": [vp, vp],
            "lvs_ndt_set_target": [vp, vp, sz, sz, i32], "lvs_ndt_set_source": [vp, vp, sz, sz, i32], "lvs_ndt_align": [vp, vp, vp],
            "lvs_ndt_get_aligned_cloud": [vp, vp, i32], "lvs_ndt_get_trace": [vp, vp, i32, vp],
            "lvs_ndt_eval_derivatives": [vp, vp, vp, i32, vp, vp, vp], "lvs_ndt_eval_hessian": [vp, vp, vp, vp],
            "lvs_ndt_calculate_score": [vp, vp, vp], "lvs_ndt_get_grid": [vp, vp, vp, vp], "lvs_ndt_num_cells": [vp, vp],
            "lvs_ndt_get_cells": [vp] * 8, "lvs_ndt_get_cell_horizontal": [vp, vp], "lvs_ndt_lookup_keys": [vp, vp, vp], "lvs_ndt_handle_batch": [vp, vp],
            "lvs_ndt_batch_create": [vp, i32, vp, i32, i32, vp], "lvs_ndt_batch_destroy": [vp],
            "lvs_ndt_batch_set_target": [vp, i32, vp, sz, sz, i32], "lvs_ndt_batch_set_source": [vp, i32, vp, sz, sz, i32],
            "lvs_ndt_batch_align": [vp, i32, vp, vp, vp, vp], "lvs_ndt_batch_last_stats": [vp, vp, vp, vp, vp],
            "lvs_ndt_batch_set_profiling": [vp, i32], "lvs_ndt_batch_set_tuning": [vp, i32, i32, i32],
            "lvs_pgo_create": [i32, i32, vp, vp], "lvs_pgo_destroy": [vp], "lvs_pgo_set_graph": [vp, i32, vp, vp, i32, vp, vp, vp, vp],
            "lvs_pgo_set_graph_typed": [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp], "lvs_pgo_set_floor_plane": [vp, vp],
            "lvs_pgo_set_poses": [vp, vp], "lvs_pgo_optimize": [vp, i32, vp], "lvs_pgo_get_poses": [vp, vp], "lvs_pgo_get_trace": [vp, vp, i32, vp],
            "lvs_pgo_set_solver_options": [vp, ctypes.c_double, i32], "lvs_pgo_compute_errors": [vp, vp, vp, vp],
            "lvs_pgo_system_size": [vp, vp, vp], "lvs_pgo_linearize": [vp, vp, vp, vp, vp],
            "lvs_pgo_solve": [vp, ctypes.c_double, ctypes.c_double, i32, vp, vp],
            "lvs_ndt_fitness_score": [vp, vp, ctypes.c_double, vp, vp], "lvs_ndt_batch_fitness_score": [vp, i32, i32, vp, ctypes.c_double, vp, vp],
            "lvs_ndt_batch_align_begin": [vp, i32, vp, vp, vp], "lvs_ndt_batch_align_end": [vp, vp],
            "lvs_prefilter_create": [i32, vp, vp], "lvs_prefilter_destroy": [vp],
            "lvs_prefilter_run": [vp, vp, sz, sz, i32, i32, ctypes.c_double, ctypes.c_double, i32, ctypes.c_float, vp, sz, i32, vp, vp],
            "lvs_prefilter_accumulate_begin": [vp], "lvs_prefilter_accumulate_add": [vp, vp, sz, sz, i32, i32, vp],
            "lvs_prefilter_accumulate_flush": [vp, i32, ctypes.c_float, vp, sz, i32, vp, vp],
            "lvs_pgo_chol_analyze": [i32, i32, vp, vp, vp], "lvs_pgo_chol_info": [vp, vp],
            "lvs_ndt_batch_total_launches": [vp, vp], "lvs_ndt_batch_transfer_bytes": [vp, vp, vp], "lvs_ndt_batch_num_cells": [vp, i32, vp, vp],
            "lvs_ndt_batch_wait_uploads": [vp], "lvs_ndt_batch_shard_init": [vp, i32, i32, i32, vp], "lvs_ndt_batch_shard_connect": [vp, vp],
            "lvs_ndt_batch_set_targets": [vp, i32, vp, vp, vp, ctypes.c_size_t, i32], "lvs_ndt_batch_set_sources": [vp, i32, vp, vp, vp, ctypes.c_size_t, i32],
        }.items():
            f = getattr(L, name)
            f.restype = i32
            f.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise LvsError(status, lib().lvs_last_error().decode("utf-8", "replace"))

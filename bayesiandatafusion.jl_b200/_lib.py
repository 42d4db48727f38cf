"""ctypes binding of libbdf_b200.so (include/bdf_b200.h). There is NO fallback: if the library is missing or fails
to load, importing the engine raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BDF_B200_LIB") or os.path.join(HERE, "libbdf_b200.so")

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
H = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/bdf_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "bdf_version": (C.c_int, []),
    "bdf_last_error": (C.c_char_p, [H]),
    "bdf_create": (C.c_int, [C.POINTER(H), C.c_int, C.c_int, C.c_int, C.c_int]),
    "bdf_destroy": (C.c_int, [H]),
    "bdf_set_stream": (C.c_int, [H, C.c_void_p]),
    "bdf_set_seed": (C.c_int, [H, C.c_uint64]),
    "bdf_add_entity": (C.c_int, [H, C.c_int64]),
    "bdf_add_relation": (C.c_int, [H, C.c_int, c_ip, C.c_int64, c_i64p, c_dp]),
    "bdf_set_relation_params": (C.c_int, [H, C.c_int, C.c_double, C.c_double]),
    "bdf_set_factors": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_get_factors": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_factors_dev": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p), c_i64p, c_i64p]),
    "bdf_ipc_export": (C.c_int, [H, C.c_int, C.c_char_p]),
    "bdf_ipc_import": (C.c_int, [H, C.c_int, C.c_int, C.c_char_p]),
    "bdf_set_features_dense": (C.c_int, [H, C.c_int, C.c_int64, C.c_int64, c_dp]),
    "bdf_compute_ff": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_set_use_ff": (C.c_int, [H, C.c_int, C.c_int]),
    "bdf_solve_full": (C.c_int, [H, C.c_int, c_dp, C.c_int, C.c_double, c_dp]),
    "bdf_step_nw_draw_on": (C.c_int, [H, C.c_int, C.c_void_p]),
    "bdf_add_entity_partitioned": (C.c_int, [H, C.c_int64, c_i32p]),
    "bdf_predict_all": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_set_relation_features": (C.c_int, [H, C.c_int, C.c_int64, C.c_int64, c_dp]),
    "bdf_sample_beta_rel": (C.c_int, [H, C.c_int, C.c_double, c_dp, c_dp, c_dp]),
    "bdf_get_relation_beta": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_set_relation_beta": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_predict_f": (C.c_int, [H, C.c_int, C.c_int64, c_i64p, c_dp, c_dp]),
    "bdf_train_sse": (C.c_int, [H, C.c_int, c_dp, c_i64p]),
    "bdf_sample_alpha": (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_dp]),
    "bdf_sample_mode": (C.c_int, [H, C.c_int, c_dp, C.c_int64, c_dp, c_dp]),
    "bdf_nw_stats": (C.c_int, [H, C.c_int, c_dp, c_dp, c_dp]),
    "bdf_set_nw_stats": (C.c_int, [H, C.c_int, C.c_double, c_dp, c_dp]),
    "bdf_stats_dev": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p), c_i64p]),
    "bdf_nw_sample": (C.c_int, [H, C.c_int, c_dp, C.c_double, c_dp, C.c_double, c_dp, c_dp, c_dp, c_dp]),
    "bdf_step_sample": (C.c_int, [H, C.c_int]),
    "bdf_step_nw_stats": (C.c_int, [H, C.c_int]),
    "bdf_step_nw_draw": (C.c_int, [H, C.c_int]),
    "bdf_sweep": (C.c_int, [H, C.c_int]),
    "bdf_advance_sweep": (C.c_int, [H]),
    "bdf_get_hyper": (C.c_int, [H, C.c_int, c_dp, c_dp]),
    "bdf_set_hyper": (C.c_int, [H, C.c_int, c_dp, c_dp]),
    "bdf_debug_row_noise": (C.c_int, [H, C.c_int, C.c_uint64, c_dp]),
    "bdf_debug_phase_clocks": (C.c_int, [H, C.c_int, c_dp, c_i64p]),
    "bdf_sweep_counter": (C.c_int64, [H]),
    "bdf_synchronize": (C.c_int, [H]),
    "bdf_launch_count": (C.c_int64, [H]),
    "bdf_predict": (C.c_int, [H, C.c_int, C.c_int64, c_i64p, c_dp]),
    "bdf_ipc_export_beta": (C.c_int, [H, C.c_int, C.c_char_p]),
    "bdf_ipc_import_beta": (C.c_int, [H, C.c_int, C.c_int, C.c_char_p]),
    "bdf_set_test": (C.c_int, [H, C.c_int, C.c_int64, c_i64p, c_dp, c_dp, C.c_double]),
    "bdf_test_reset": (C.c_int, [H, C.c_int]),
    "bdf_predict_accumulate": (C.c_int, [H, C.c_int, C.c_int, C.c_double, C.c_double, c_dp]),
    "bdf_get_test_predictions": (C.c_int, [H, C.c_int, c_dp, c_dp, c_dp]),
    "bdf_set_async": (C.c_int, [H, C.c_int]),
    "bdf_nw_sample_async": (C.c_int, [H, C.c_int, c_dp, C.c_double, c_dp, C.c_double, c_dp, c_dp]),
    "bdf_nw_sample_fetch": (C.c_int, [H, C.c_int, c_dp, c_dp]),
    "bdf_set_features_sbm": (C.c_int, [H, C.c_int, C.c_int64, C.c_int64, C.c_int64, c_i32p, c_i32p]),
    "bdf_set_features_csc": (C.c_int, [H, C.c_int, C.c_int64, C.c_int64, c_i64p, c_i64p, c_dp]),
    "bdf_debug_features_csr": (C.c_int, [H, C.c_int, C.c_int, c_i32p, c_i32p]),
    "bdf_spmm": (C.c_int, [H, C.c_int, C.c_int, c_dp, C.c_int, c_dp]),
    "bdf_ata_mul": (C.c_int, [H, C.c_int, c_dp, C.c_double, c_dp]),
    "bdf_cg_solve": (C.c_int, [H, C.c_int, c_dp, C.c_int, C.c_double, C.c_double, C.c_int64, c_dp, c_ip]),
    "bdf_set_beta": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_get_beta": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_update_uhat": (C.c_int, [H, C.c_int, c_dp, c_dp]),
    "bdf_sample_mode_uhat": (C.c_int, [H, C.c_int, c_dp, c_dp]),
    "bdf_nw_stats_uhat": (C.c_int, [H, C.c_int, c_dp, c_dp, c_dp]),
    "bdf_beta_gram": (C.c_int, [H, C.c_int, c_dp]),
    "bdf_sample_beta": (C.c_int, [H, C.c_int, c_dp, c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "bdf_debug_ata_time": (C.c_int, [H, C.c_int, C.c_int, c_dp]),
    "bdf_debug_ata_time_window": (C.c_int, [H, C.c_int, C.c_int, C.c_int, c_dp]),
    "bdf_sample_lambda_beta": (C.c_int, [H, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, c_dp, c_dp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libbdf_b200.so and bind every entry point. Raises if it is absent: the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found — build it with `python bayesiandatafusion.jl_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

"""ctypes binding of libhygeia_b200.so (include/hygeia_b200.h).  Fails loudly: no CUDA extension -> no product."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HYGEIA_B200_LIB") or os.path.join(HERE, "csrc", "libhygeia_b200.so")   # override: kernel-variant experiments

# every symbol include/hygeia_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "hyg_create", "hyg_destroy", "hyg_last_error", "hyg_create_error", "hyg_version", "hyg_stream",
    "hyg_sg_set_model", "hyg_sg_set_vartheta", "hyg_sg_set_theta", "hyg_sg_get_tables",
    "hyg_sg_add_dataset", "hyg_sg_clear", "hyg_sg_default_run_args", "hyg_sg_set_chains",
    "hyg_sg_set_segmentation", "hyg_sg_filter_units", "hyg_sg_set_zero_copy_outputs",
    "hyg_sg_emission", "hyg_sg_filter", "hyg_sg_download", "hyg_sg_device_outputs", "hyg_sync", "hyg_sg_timings", "hyg_sg_get_logobs",
    "hyg_sg_run_online_combined_inference", "hyg_sg_sample_theta_prior", "hyg_philox_uniform",
    "hyg_tg_set_model", "hyg_tg_run", "hyg_tg_hazard_table", "hyg_tg_reference_hazard_table",
    "hyg_tg_site_statistics", "hyg_fdr_procedure", "hyg_weighted_fdr_procedure",
]


class HygChain(C.Structure):
    _fields_ = [
        ("dataset", C.c_uint32), ("seed", C.c_uint64), ("chain_id", C.c_uint32),
        ("uniforms", C.c_void_p), ("positions", C.c_void_p),
        ("regime_probs", C.c_void_p), ("logz", C.c_void_p), ("theta_trace", C.c_void_p),
        ("k_kept", C.c_void_p), ("drew_uniform", C.c_void_p), ("n_pending", C.c_void_p), ("n_curr", C.c_void_p),
        ("finalised_at", C.c_void_p), ("support_hash", C.c_void_p), ("tie_flags", C.c_void_p), ("status", C.c_int32 * 8),
        ("overlap_max_abs", C.c_double),
    ]


class HygRunArgs(C.Structure):
    _fields_ = [
        ("n_particles_max", C.c_uint32), ("smc_proposal_type", C.c_uint32), ("smc_resample_type", C.c_uint32),
        ("use_online_marginal_smoothing", C.c_int32), ("epsilon", C.c_double),
        ("use_online_parameter_estimation", C.c_int32), ("normalise_gradients", C.c_int32), ("use_adam", C.c_int32),
        ("n_steps_without_parameter_update", C.c_uint32), ("learning_rate_exponent", C.c_double),
        ("learning_rate_factor", C.c_double), ("lag_capacity", C.c_uint32), ("allow_forced_emission", C.c_int32),
        ("resample_full_sort", C.c_int32),
    ]


class HygTgModel(C.Structure):
    _fields_ = [
        ("R", C.c_uint32), ("minimum_duration", C.c_uint32), ("num_resampled", C.c_uint32), ("num_backward", C.c_uint32),
        ("log_p_control", C.c_void_p), ("omega_control", C.c_void_p), ("omega_case", C.c_void_p),
        ("kappa_control", C.c_void_p), ("kappa_case", C.c_void_p),
        ("merge_prob", C.c_double), ("split_prob", C.c_double),
        ("rho_control", C.c_void_p), ("rho_case", C.c_void_p), ("d_max", C.c_uint32), ("hazard_mode", C.c_uint32), ("sort_preselect", C.c_uint32 * 2), ("sort_scratch_from", C.c_uint32),
    ]


class HygTgChain(C.Structure):
    _fields_ = [
        ("control_dataset", C.c_uint32), ("case_dataset", C.c_uint32), ("seed", C.c_uint64), ("chain_id", C.c_uint32),
        ("trajectories", C.c_void_p), ("log_normalizing_constant", C.c_void_p), ("taps", C.c_void_p),
    ]


_lib = None


def load():
    """Load the CUDA library or raise (there is no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  hygeia_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.hyg_create.restype = C.c_void_p
    lib.hyg_create.argtypes = [C.c_int]
    lib.hyg_destroy.argtypes = [C.c_void_p]
    lib.hyg_last_error.restype = C.c_char_p
    lib.hyg_last_error.argtypes = [C.c_void_p]
    lib.hyg_create_error.restype = C.c_char_p
    lib.hyg_version.restype = C.c_char_p
    lib.hyg_stream.restype = C.c_void_p
    lib.hyg_stream.argtypes = [C.c_void_p]
    lib.hyg_sg_set_model.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.hyg_sg_set_vartheta.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.hyg_sg_set_theta.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
    lib.hyg_sg_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.hyg_sg_add_dataset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64]
    lib.hyg_sg_clear.argtypes = [C.c_void_p]
    lib.hyg_sg_default_run_args.argtypes = [C.POINTER(HygRunArgs)]
    lib.hyg_sg_set_chains.argtypes = [C.c_void_p, C.POINTER(HygChain), C.c_uint32]
    lib.hyg_sg_set_segmentation.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.hyg_sg_filter_units.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.hyg_sg_set_zero_copy_outputs.argtypes = [C.c_void_p, C.c_int]
    lib.hyg_sg_emission.argtypes = [C.c_void_p]
    lib.hyg_sg_filter.argtypes = [C.c_void_p, C.POINTER(HygRunArgs)]
    lib.hyg_sg_download.argtypes = [C.c_void_p, C.POINTER(HygChain), C.c_uint32]
    lib.hyg_sg_device_outputs.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.hyg_sync.argtypes = [C.c_void_p]
    lib.hyg_sg_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.hyg_sg_get_logobs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.hyg_sg_run_online_combined_inference.argtypes = [
        C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
        C.POINTER(HygRunArgs), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
    lib.hyg_sg_sample_theta_prior.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
    lib.hyg_tg_set_model.argtypes = [C.c_void_p, C.POINTER(HygTgModel), C.c_uint64]
    lib.hyg_tg_run.argtypes = [C.c_void_p, C.POINTER(HygTgChain), C.c_uint32, C.POINTER(C.c_float)]
    lib.hyg_tg_hazard_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.hyg_tg_reference_hazard_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.hyg_tg_site_statistics.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.hyg_fdr_procedure.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]
    lib.hyg_weighted_fdr_procedure.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                               C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_double)]
    lib.hyg_philox_uniform.restype = C.c_double
    lib.hyg_philox_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    _lib = lib
    return lib

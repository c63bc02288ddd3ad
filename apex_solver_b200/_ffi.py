"""ctypes view of include/apex_gpu.h.

The structures are declared once here; `bind(lib, prefix)` attaches argument/return types to a
loaded shared library whose symbols are `<prefix>ctx_create`, ... . The product library is
`csrc/libapex_gpu.so` (prefix ``apex_``). The test oracle exports the same entry points with the
prefix ``oracle_`` and is bound by the tests, never by this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libapex_gpu.so")

# ---- status codes (include/apex_gpu.h) -------------------------------------------------------
OK = 0
ERR_FACTORIZATION_FAILED = -1
ERR_SINGULAR_MATRIX = -2
ERR_INVALID_INPUT = -5
ERR_INVALID_STATE = -6
ERR_LINEAR_SOLVE_FAILED = -10
ERR_INVALID_PARAMETERS = -11
ERR_NUMERICAL_INSTABILITY = -12
ERR_EMPTY_PROBLEM = -13
ERR_NO_RESIDUAL_BLOCKS = -14
ERR_CUDA = -20
ERR_NCCL = -21
ERR_NO_DEVICE = -22
ERR_UNSUPPORTED = -23
ERR_IO = -30
ERR_PARSE = -31
ERR_INVALID_NUMBER = -32
ERR_MISSING_FIELDS = -33

ERROR_NAMES = {
    ERR_FACTORIZATION_FAILED: "LinAlgError::FactorizationFailed",
    ERR_SINGULAR_MATRIX: "LinAlgError::SingularMatrix",
    ERR_INVALID_INPUT: "InvalidInput",
    ERR_INVALID_STATE: "LinAlgError::InvalidState",
    ERR_LINEAR_SOLVE_FAILED: "OptimizerError::LinearSolveFailed",
    ERR_INVALID_PARAMETERS: "OptimizerError::InvalidParameters",
    ERR_NUMERICAL_INSTABILITY: "OptimizerError::NumericalInstability",
    ERR_EMPTY_PROBLEM: "OptimizerError::EmptyProblem",
    ERR_NO_RESIDUAL_BLOCKS: "OptimizerError::NoResidualBlocks",
    ERR_CUDA: "CUDA error",
    ERR_NCCL: "NCCL error",
    ERR_NO_DEVICE: "no CUDA device (there is no CPU fallback)",
    ERR_UNSUPPORTED: "unsupported on the GPU path",
    ERR_IO: "IoError::Io",
    ERR_PARSE: "IoError::Parse",
    ERR_INVALID_NUMBER: "IoError::InvalidNumber",
    ERR_MISSING_FIELDS: "IoError::MissingFields",
}

# ---- enums -----------------------------------------------------------------------------------
CAM_BAL, CAM_PINHOLE, CAM_KANNALA_BRANDT, CAM_DOUBLE_SPHERE, CAM_RADTAN, CAM_UCM, CAM_EUCM, CAM_FOV, CAM_FTHETA = range(9)
CAM_INTR_DIM = {CAM_BAL: 3, CAM_PINHOLE: 4, CAM_KANNALA_BRANDT: 8, CAM_DOUBLE_SPHERE: 6, CAM_RADTAN: 9,
                CAM_UCM: 5, CAM_EUCM: 6, CAM_FOV: 5, CAM_FTHETA: 6}
OPT_POSE, OPT_LANDMARK, OPT_INTRINSIC, OPT_SHARED_INTRINSICS = 1, 2, 4, 8
(LOSS_NONE, LOSS_L2, LOSS_L1, LOSS_HUBER, LOSS_CAUCHY, LOSS_FAIR, LOSS_GEMAN_MCCLURE, LOSS_WELSCH, LOSS_TUKEY,
 LOSS_ANDREWS, LOSS_RAMSAY_EA, LOSS_TRIMMED_MEAN, LOSS_LP_NORM, LOSS_BARRON, LOSS_T_DISTRIBUTION) = range(15)
SCHUR_EXPLICIT, SCHUR_IMPLICIT, SCHUR_EXPLICIT_PCG = 0, 1, 2
PRECOND_NONE, PRECOND_BLOCK_DIAGONAL, PRECOND_SCHUR_JACOBI = 0, 1, 2


class ApexError(RuntimeError):
    """Raised for a negative apex_status; `.status` holds the code (mirrors the reference's Err(..))."""

    def __init__(self, status: int, message: str = ""):
        self.status = int(status)
        name = ERROR_NAMES.get(self.status, f"status {status}")
        super().__init__(f"{name}: {message}" if message else name)


# ---- PODs ------------------------------------------------------------------------------------
class CtxDesc(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("reserved", C.c_int32),
                ("nccl_unique_id", C.c_void_p)]


class LossSpec(C.Structure):
    _fields_ = [("loss_id", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 4)]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("camera_model", C.c_int32), ("opt_flags", C.c_uint32), ("intr_dim", C.c_int32), ("intr_vars_present", C.c_int32),
        ("ncam", C.c_uint32), ("npts", C.c_uint32), ("nobs", C.c_uint64),
        ("pose", C.c_void_p), ("intr", C.c_void_p), ("pt", C.c_void_p),
        ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_uv", C.c_void_p),
        ("loss_id", C.c_int32), ("reserved0", C.c_int32), ("loss_params", C.c_double * 4),
        ("pose_fixed", C.c_void_p), ("intr_fixed", C.c_void_p), ("pt_fixed", C.c_void_p),
        ("obs_loss", C.c_void_p), ("loss_table", C.c_void_p), ("n_losses", C.c_int32), ("reserved1", C.c_int32),
    ]


class LmConfig(C.Structure):
    _fields_ = [
        ("schur_variant", C.c_int32), ("schur_preconditioner", C.c_int32), ("max_iterations", C.c_int32), ("cg_max_iterations", C.c_int32),
        ("cost_tolerance", C.c_double), ("parameter_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("timeout_seconds", C.c_double),
        ("damping", C.c_double), ("damping_min", C.c_double), ("damping_max", C.c_double),
        ("damping_increase_factor", C.c_double), ("damping_decrease_factor", C.c_double), ("damping_nu", C.c_double),
        ("trust_region_radius", C.c_double), ("min_step_quality", C.c_double), ("good_step_quality", C.c_double),
        ("min_diagonal", C.c_double), ("max_diagonal", C.c_double), ("min_cost_threshold", C.c_double),
        ("min_trust_region_radius", C.c_double), ("max_condition_number", C.c_double), ("min_relative_decrease", C.c_double),
        ("cg_tolerance", C.c_double), ("use_jacobi_scaling", C.c_int32), ("compute_covariances", C.c_int32),
    ]


class LmResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("iterations", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("elapsed_seconds", C.c_double), ("final_gradient_norm", C.c_double), ("final_parameter_update_norm", C.c_double),
        ("cost_evaluations", C.c_int32), ("jacobian_evaluations", C.c_int32), ("successful_steps", C.c_int32),
        ("unsuccessful_steps", C.c_int32), ("final_damping", C.c_double), ("final_damping_nu", C.c_double),
        ("linear_iterations", C.c_int64),
    ]


class IterTrace(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("accepted", C.c_int32), ("ls_iter", C.c_int32), ("reserved", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_norm", C.c_double), ("step_norm", C.c_double),
        ("tr_ratio", C.c_double), ("tr_radius", C.c_double), ("new_cost", C.c_double), ("predicted_reduction", C.c_double),
        ("parameter_norm", C.c_double), ("iter_time_ms", C.c_double),
    ]


class Profile(C.Structure):
    _fields_ = [("lm_device_ms", C.c_double), ("matvec_ms", C.c_double), ("matvec_launches", C.c_int64),
                ("linearize_ms", C.c_double), ("linearize_launches", C.c_int64),
                ("schur_form_ms", C.c_double), ("schur_forms", C.c_int64), ("cholesky_ms", C.c_double), ("cholesky_factorizations", C.c_int64),
                ("cholesky_n", C.c_uint64), ("upload_h2d_bytes", C.c_uint64)]


class LayoutStats(C.Structure):
    _fields_ = [("shard_block", C.c_uint32), ("npts_local", C.c_uint32), ("nobs_local", C.c_uint64), ("ntiles", C.c_uint32), ("nlong_tiles", C.c_uint32),
                ("nchunks", C.c_uint32), ("nnormal_chunks", C.c_uint32), ("ncam_items", C.c_uint32), ("max_segments_per_chunk", C.c_uint32),
                ("nsegments", C.c_uint64), ("slots_used", C.c_uint64), ("consistent", C.c_int32), ("reserved", C.c_int32), ("build_ms", C.c_double),
                ("mv_ranges", C.c_uint32), ("mv_window", C.c_uint32), ("mv_nwindows", C.c_uint32), ("reserved2", C.c_uint32), ("mv_rows", C.c_uint64)]


class BalView(C.Structure):
    _fields_ = [("ncam", C.c_uint32), ("npts", C.c_uint32), ("nobs", C.c_uint64), ("cameras", C.c_void_p), ("points", C.c_void_p),
                ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_uv", C.c_void_p)]


class Dims(C.Structure):
    _fields_ = [("ncam", C.c_uint32), ("npts", C.c_uint32), ("nobs", C.c_uint64), ("intr_dim", C.c_int32), ("dc", C.c_int32),
                ("cam_dof", C.c_uint64), ("lm_dof", C.c_uint64), ("npts_local", C.c_uint32), ("flags", C.c_uint32),
                ("nobs_local", C.c_uint64)]


class ObserverMetrics(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("accepted", C.c_int32), ("cost", C.c_double), ("gradient_norm", C.c_double),
                ("damping", C.c_double), ("step_norm", C.c_double), ("step_quality", C.c_double)]


ON_STEP = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(ObserverMetrics))
ON_COMPLETE = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32)


class Observer(C.Structure):
    _fields_ = [("on_step", ON_STEP), ("on_optimization_complete", ON_COMPLETE), ("user", C.c_void_p)]


# Every symbol include/apex_gpu.h declares (without prefix): name -> (restype, argtypes)
P = C.POINTER
_DBL = C.c_void_p  # double* passed as raw addresses of numpy buffers
SYMBOLS = {
    "abi_version": (C.c_int32, []),
    "device_count": (C.c_int32, []),
    "lm_config_default": (None, [P(LmConfig)]),
    "lm_config_for_bundle_adjustment": (None, [P(LmConfig)]),
    "nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "ctx_create": (C.c_int32, [P(CtxDesc), P(C.c_void_p)]),
    "ctx_destroy": (None, [C.c_void_p]),
    "last_error": (C.c_char_p, [C.c_void_p]),
    "problem_upload": (C.c_int32, [C.c_void_p, P(ProblemDesc)]),
    "get_dims": (C.c_int32, [C.c_void_p, P(Dims)]),
    "params_upload": (C.c_int32, [C.c_void_p, _DBL, _DBL, _DBL]),
    "params_download": (C.c_int32, [C.c_void_p, _DBL, _DBL, _DBL]),
    "linearize": (C.c_int32, [C.c_void_p, C.c_double]),
    "cost": (C.c_int32, [C.c_void_p, P(C.c_double)]),
    "get_linearization": (C.c_int32, [C.c_void_p, _DBL, _DBL, _DBL]),
    "get_blocks": (C.c_int32, [C.c_void_p, _DBL, _DBL, _DBL, _DBL, _DBL]),
    "schur_matvec": (C.c_int32, [C.c_void_p, _DBL, _DBL]),
    "schur_matvec_bench": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, P(C.c_double)]),
    "solve_augmented": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, _DBL, _DBL,
                                    P(C.c_double), P(C.c_int32)]),
    "get_step": (C.c_int32, [C.c_void_p, _DBL, _DBL]),
    "lm_solve": (C.c_int32, [C.c_void_p, P(LmConfig), P(LmResult), P(IterTrace), C.c_int32]),
    "add_observer": (C.c_int32, [C.c_void_p, P(Observer)]),
    "clear_observers": (C.c_int32, [C.c_void_p]),
    "kernel_launches": (C.c_int64, [C.c_void_p]),
    "profile_enable": (C.c_int32, [C.c_void_p, C.c_int32]),
    "profile_read": (C.c_int32, [C.c_void_p, P(Profile)]),
    "layout_stats_compute": (C.c_int32, [P(ProblemDesc), C.c_int32, C.c_int32, P(LayoutStats)]),
    "dense_cholesky_bench": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_int32, P(C.c_double)]),
    "bal_load": (C.c_int32, [C.c_char_p, P(C.c_void_p)]),
    "bal_from_arrays": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_void_p)]),
    "bal_view_get": (C.c_int32, [C.c_void_p, P(BalView)]),
    "bal_write": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "bal_build_problem": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_int32, P(ProblemDesc)]),
    "bal_free": (None, [C.c_void_p]),
    "bal_last_error": (C.c_char_p, []),
    "shard_info": (C.c_int32, [C.c_uint32, C.c_uint64, C.c_void_p, C.c_int32, C.c_int32, P(C.c_uint32), P(C.c_uint32), P(C.c_uint64)]),
}
# entry points the oracle does not implement (GPU-only plumbing)
GPU_ONLY = {"device_count", "nccl_unique_id", "schur_matvec_bench", "kernel_launches", "profile_enable", "profile_read", "shard_info", "layout_stats_compute", "dense_cholesky_bench",
            "bal_load", "bal_from_arrays", "bal_view_get", "bal_write", "bal_build_problem", "bal_free", "bal_last_error"}


def bind(lib: C.CDLL, prefix: str, skip=()) -> None:
    for name, (res, args) in SYMBOLS.items():
        if name in skip:
            continue
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args


_lib = None


def load_library() -> C.CDLL:
    """Load csrc/libapex_gpu.so. Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C apex_solver_b200/csrc`. The GPU path has no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        bind(lib, "apex_")
        _lib = lib
    return _lib


def ptr(a):
    """Raw address of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def as_f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)

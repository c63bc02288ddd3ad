"""Loads the TEST ORACLE (oracle/liboracle.so, CPU restatement of the reference) behind the same Python
`Context` wrapper the GPU library uses. Only tests/, smoke() and bench.py's cpu_baseline leg use this."""
import ctypes as C
import os
import subprocess

from apex_solver_b200 import _ffi as F
from apex_solver_b200.context import Context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
_lib = None


def build_oracle():
    src = os.path.join(ORACLE_DIR, "apex_oracle.cpp")
    if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def oracle_lib():
    global _lib
    if _lib is None:
        build_oracle()
        lib = C.CDLL(ORACLE_LIB, mode=C.RTLD_LOCAL)
        F.bind(lib, "oracle_", skip=F.GPU_ONLY)
        # the oracle's get_blocks takes one more argument: which flavour of guarded inverse to return
        lib.oracle_get_blocks.argtypes = F.SYMBOLS["get_blocks"][1] + [C.c_int32]
        lib.oracle_schur_matvec_partial.restype = C.c_int32
        lib.oracle_schur_matvec_partial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32]
        lib.oracle_num_threads.restype = C.c_int32
        lib.oracle_compute_cost.restype = C.c_double
        lib.oracle_compute_cost.argtypes = [C.c_void_p, C.c_uint64]
        lib.oracle_compute_step_quality.restype = C.c_double
        lib.oracle_compute_step_quality.argtypes = [C.c_double] * 3
        lib.oracle_update_damping.restype = C.c_int32
        lib.oracle_update_damping.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
        lib.oracle_check_convergence.restype = C.c_int32
        lib.oracle_check_convergence.argtypes = [C.c_int32] + [C.c_double] * 6 + [C.c_int32, C.c_int32] + [C.c_double] * 7
        lib.oracle_loss_evaluate.argtypes = [C.c_int32, C.c_void_p, C.c_double, C.c_void_p]
        lib.oracle_corrector.argtypes = [C.c_int32, C.c_void_p, C.c_double, C.c_void_p]
        VP, I32, U32, DBL = C.c_void_p, C.c_int32, C.c_uint32, C.c_double
        sigs = {
            "oracle_project": (I32, [I32, VP, VP, VP]),
            "oracle_jacobian_point": (None, [I32, VP, VP, VP]),
            "oracle_jacobian_intrinsics": (None, [I32, VP, VP, VP]),
            "oracle_camera_intr_dim": (I32, [I32]),
            "oracle_se3_normalize": (None, [VP, VP]),
            "oracle_se3_act": (None, [VP, VP, VP]),
            "oracle_se3_plus": (None, [VP, VP, VP]),
            "oracle_rotation_matrix": (None, [VP, VP]),
            "oracle_linearize_block": (None, [I32, U32, I32, VP, VP, VP, VP, VP, VP, VP]),
            "oracle_invert_landmark_block": (I32, [VP, DBL, I32, VP]),
            "oracle_inverse_n": (I32, [I32, VP, VP]),
            "oracle_schur_complement_dense": (None, [U32, U32, VP, VP, VP, VP]),
            "oracle_reduced_gradient_dense": (None, [U32, U32, VP, VP, VP, VP, VP]),
            "oracle_back_substitute_dense": (None, [U32, U32, VP, VP, VP, VP, VP]),
            "oracle_solve_cholesky_dense": (I32, [U32, VP, VP, VP]),
            "oracle_solve_pcg_dense": (I32, [U32, VP, VP, VP, I32, DBL]),
            "oracle_get_column_layout": (I32, [VP, VP, VP, VP]),
            "oracle_set_num_threads": (None, [I32]),
            "oracle_set_reverse_order": (None, [I32]),
            "oracle_set_parallel_sweeps": (None, [I32]),
            "oracle_apply_step": (I32, [VP, VP, VP, DBL]),
            "oracle_back_substitute": (I32, [VP, VP, I32, VP]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class OracleContext(Context):
    def __init__(self):
        super().__init__(oracle_lib(), "oracle_")

    def get_blocks(self, implicit_flavour=True):
        return super().get_blocks(1 if implicit_flavour else 0)

    def apply_step(self, step_cam, step_pt, sign=1.0):
        sc, sp = F.as_f64(step_cam), F.as_f64(step_pt)
        self._check(self._lib.oracle_apply_step(self._h, F.ptr(sc), F.ptr(sp), float(sign)))

    def back_substitute(self, step_cam, implicit_flavour=True):
        import numpy as np
        sc = F.as_f64(step_cam)
        sp = np.empty((self.problem.npts, 3))
        self._check(self._lib.oracle_back_substitute(self._h, F.ptr(sc), 1 if implicit_flavour else 0, F.ptr(sp)))
        return sp

    def schur_matvec_partial(self, x, p0, p1, add_hcc):
        import numpy as np
        p = self.problem
        x = F.as_f64(x).reshape(p.ncam * p.dc)
        y = np.empty_like(x)
        self._check(self._lib.oracle_schur_matvec_partial(self._h, F.ptr(x), F.ptr(y), int(p0), int(p1), 1 if add_hcc else 0))
        return y


class OracleVariant(OracleContext):
    """The oracle in another SUMMATION ORDER: same arithmetic per observation / landmark, block sums and landmark sweeps
    accumulated in reverse order (`reverse`) and / or spread over `threads` OpenMP threads with thread-private partial
    results added in thread order (`parallel`; the reference runs these sweeps on one thread, any rayon thread count
    reorders its other sums the same way). oracle-vs-OracleVariant is the rounding yardstick of the parity tests: the
    reference algorithm itself is only reproducible down to that floor. Also the CPU stand-in "device under test" that
    exercises the parity harness without a GPU, and (parallel=True) the fast CPU arm of the full-size tests / bench.py."""

    def __init__(self, reverse=False, parallel=False, threads=None):
        super().__init__()
        self._flags = (1 if reverse else 0, 1 if parallel else 0, threads)

    def _wrap(name):
        def call(self, *a, **kw):
            lib = oracle_lib()
            rev, par, threads = self._flags
            before = lib.oracle_num_threads()
            lib.oracle_set_reverse_order(rev)
            lib.oracle_set_parallel_sweeps(par)
            if threads:
                lib.oracle_set_num_threads(int(threads))
            try:
                return getattr(OracleContext, name)(self, *a, **kw)
            finally:
                lib.oracle_set_reverse_order(0)
                lib.oracle_set_parallel_sweeps(0)
                if threads:
                    lib.oracle_set_num_threads(before)
        return call

    get_step = OracleContext.get_step
    linearize = _wrap("linearize")
    cost = _wrap("cost")
    schur_matvec = _wrap("schur_matvec")
    solve_augmented = _wrap("solve_augmented")
    lm_solve = _wrap("lm_solve")
    del _wrap

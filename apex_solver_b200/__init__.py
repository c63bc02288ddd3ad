"""apex_solver_b200 — B200-native (sm_100a) bundle-adjustment Levenberg–Marquardt path of apex-solver.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/apex_gpu.h), a ctypes view of
that ABI, the host-side mirror of the reference's Problem / Factor / LossFunction / LevenbergMarquardt
surface, the BAL file format and the synthetic BAL-shaped generator. There is no CPU fallback.
"""
from . import _ffi  # noqa: F401
from ._ffi import ApexError  # noqa: F401
from .context import BAProblem, Context, GpuContext  # noqa: F401

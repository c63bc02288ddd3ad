"""BAL files and the CLI's problem construction: Python mirror of `BalLoader` / `BalDataset`
(crates/apex-io/src/bal.rs:60-400) and of `run_bundle_adjustment` (bin/bundle_adjustment.rs:212-441) over the host-only
C entry points `apex_bal_*` (csrc/bal_io.cpp). Nothing here parses or computes: every function is one C call plus a copy
of the result into numpy arrays."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _ffi as F
from .context import BAProblem

BUNDLE_ADJUSTMENT = 0   # OptimizationType::BundleAdjustment: pose + landmarks
SELF_CALIBRATION = 1    # OptimizationType::SelfCalibration (default): pose + landmarks + intrinsics
OPTIMIZATION_TYPES = {"bundle-adjustment": BUNDLE_ADJUSTMENT, "self-calibration": SELF_CALIBRATION,
                      "only-pose": 2, "only-landmarks": 3, "only-intrinsics": 4}   # the last three are not functional in the reference


def _fail(st: int):
    raise F.ApexError(st, F.load_library().apex_bal_last_error().decode())


@dataclass
class BalDataset:
    """`BalDataset` (bal.rs:84-93) as arrays: cameras[ncam, 9] = rx ry rz tx ty tz f k1 k2 (f already normalised,
    bal.rs:107-113), points[npts, 3], observations as (camera index, point index, pixel)."""
    cameras: np.ndarray
    points: np.ndarray
    obs_cam: np.ndarray
    obs_pt: np.ndarray
    obs_uv: np.ndarray

    def __post_init__(self):
        self.cameras = np.ascontiguousarray(self.cameras, dtype=np.float64).reshape(-1, 9)
        self.points = np.ascontiguousarray(self.points, dtype=np.float64).reshape(-1, 3)
        self.obs_cam = np.ascontiguousarray(self.obs_cam, dtype=np.uint32).reshape(-1)
        self.obs_pt = np.ascontiguousarray(self.obs_pt, dtype=np.uint32).reshape(-1)
        self.obs_uv = np.ascontiguousarray(self.obs_uv, dtype=np.float64).reshape(-1, 2)

    def _handle(self) -> C.c_void_p:
        h = C.c_void_p()
        st = F.load_library().apex_bal_from_arrays(self.cameras.shape[0], self.points.shape[0], self.obs_cam.shape[0], F.ptr(self.cameras), F.ptr(self.points),
                                                   F.ptr(self.obs_cam), F.ptr(self.obs_pt), F.ptr(self.obs_uv), C.byref(h))
        if st != F.OK:
            _fail(st)
        return h

    def write(self, path: str) -> None:
        lib = F.load_library()
        h = self._handle()
        try:
            st = lib.apex_bal_write(h, os.fsencode(path))
            if st != F.OK:
                _fail(st)
        finally:
            lib.apex_bal_free(h)

    def problem(self, num_points: int | None = None, optimization_type: int | str = SELF_CALIBRATION) -> BAProblem:
        """The factor graph bin/bundle_adjustment.rs builds: `-n num_points`, `-t optimization_type`."""
        if isinstance(optimization_type, str):
            optimization_type = OPTIMIZATION_TYPES[optimization_type]
        lib = F.load_library()
        h = self._handle()
        try:
            d = F.ProblemDesc()
            n = self.points.shape[0] if num_points is None else int(num_points)
            st = lib.apex_bal_build_problem(h, n, int(optimization_type), C.byref(d))
            if st != F.OK:
                _fail(st)
            arr = lambda p, shape, dt: np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=shape).copy() if int(np.prod(shape)) else np.zeros(shape, dtype=dt)
            return BAProblem(camera_model=d.camera_model, opt_flags=d.opt_flags, pose=arr(d.pose, (d.ncam, 7), C.c_double), intr=arr(d.intr, (d.ncam, 3), C.c_double),
                             pt=arr(d.pt, (d.npts, 3), C.c_double), obs_cam=arr(d.obs_cam, (d.nobs,), C.c_uint32), obs_pt=arr(d.obs_pt, (d.nobs,), C.c_uint32),
                             obs_uv=arr(d.obs_uv, (d.nobs, 2), C.c_double), loss_id=d.loss_id, loss_params=tuple(d.loss_params),
                             intr_vars_present=bool(d.intr_vars_present), pose_fixed=arr(d.pose_fixed, (d.ncam,), C.c_uint8))
        finally:
            lib.apex_bal_free(h)


def load_bal(path: str) -> BalDataset:
    """`BalLoader::load` (bal.rs:138-200). Raises ApexError with the reference's IoError text."""
    lib = F.load_library()
    h = C.c_void_p()
    st = lib.apex_bal_load(os.fsencode(path), C.byref(h))
    if st != F.OK:
        _fail(st)
    try:
        v = F.BalView()
        lib.apex_bal_view_get(h, C.byref(v))
        arr = lambda p, shape, dt: np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=shape).copy() if int(np.prod(shape)) else np.zeros(shape, dtype=dt)
        return BalDataset(cameras=arr(v.cameras, (v.ncam, 9), C.c_double), points=arr(v.points, (v.npts, 3), C.c_double), obs_cam=arr(v.obs_cam, (v.nobs,), C.c_uint32),
                          obs_pt=arr(v.obs_pt, (v.nobs,), C.c_uint32), obs_uv=arr(v.obs_uv, (v.nobs, 2), C.c_double))
    finally:
        lib.apex_bal_free(h)


def dataset_from_problem(prob: BAProblem) -> BalDataset:
    """BAL form of a BAL-camera problem (generator output): quaternion -> axis-angle, intrinsics [f, k1, k2]."""
    from .synth import quat_to_axis_angle
    assert prob.camera_model == F.CAM_BAL
    cams = np.concatenate([quat_to_axis_angle(prob.pose[:, 3:]), prob.pose[:, :3], prob.intr], axis=1)
    return BalDataset(cameras=cams, points=prob.pt, obs_cam=prob.obs_cam, obs_pt=prob.obs_pt, obs_uv=prob.obs_uv)

"""Thin object wrapper over the C ABI of include/apex_gpu.h.

`Context(lib, prefix)` works on any library exporting that ABI; `GpuContext` binds the product
library (csrc/libapex_gpu.so). Nothing here computes: every method is one C call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _ffi as F


@dataclass
class BAProblem:
    """SoA form of the factor graph bin/bundle_adjustment.rs:212-441 builds (one ProjectionFactor with a
    single observation + one loss per residual block)."""

    camera_model: int
    opt_flags: int
    pose: np.ndarray       # [ncam,7] tx,ty,tz,qw,qx,qy,qz (world->camera)
    intr: np.ndarray       # [ncam,K]
    pt: np.ndarray         # [npts,3]
    obs_cam: np.ndarray    # [nobs] u32
    obs_pt: np.ndarray     # [nobs] u32
    obs_uv: np.ndarray     # [nobs,2]
    loss_id: int = F.LOSS_HUBER
    loss_params: tuple = (1.0, 0.0, 0.0, 0.0)
    intr_vars_present: bool = True
    pose_fixed: Optional[np.ndarray] = None   # [ncam] u8 bitmask
    intr_fixed: Optional[np.ndarray] = None   # [ncam] u16 bitmask
    pt_fixed: Optional[np.ndarray] = None     # [npts] u8 bitmask
    obs_loss: Optional[np.ndarray] = None     # [nobs] u8 index into loss_table (per-block loss functions), None = uniform loss
    loss_table: Optional[list] = None         # [(loss_id, p0, p1, ...), ...]
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.pose = np.ascontiguousarray(self.pose, dtype=np.float64).reshape(-1, 7)
        K = F.CAM_INTR_DIM[self.camera_model]
        self.intr = np.ascontiguousarray(self.intr, dtype=np.float64).reshape(-1, K)
        self.pt = np.ascontiguousarray(self.pt, dtype=np.float64).reshape(-1, 3)
        self.obs_cam = np.ascontiguousarray(self.obs_cam, dtype=np.uint32).reshape(-1)
        self.obs_pt = np.ascontiguousarray(self.obs_pt, dtype=np.uint32).reshape(-1)
        self.obs_uv = np.ascontiguousarray(self.obs_uv, dtype=np.float64).reshape(-1, 2)
        if self.pose_fixed is not None:
            self.pose_fixed = np.ascontiguousarray(self.pose_fixed, dtype=np.uint8)
        if self.intr_fixed is not None:
            self.intr_fixed = np.ascontiguousarray(self.intr_fixed, dtype=np.uint16)
        if self.pt_fixed is not None:
            self.pt_fixed = np.ascontiguousarray(self.pt_fixed, dtype=np.uint8)
        if self.obs_loss is not None:
            self.obs_loss = np.ascontiguousarray(self.obs_loss, dtype=np.uint8).reshape(-1)

    @property
    def ncam(self): return self.pose.shape[0]
    @property
    def npts(self): return self.pt.shape[0]
    @property
    def nobs(self): return self.obs_cam.shape[0]
    @property
    def K(self): return F.CAM_INTR_DIM[self.camera_model]
    @property
    def dc(self): return 6 + (self.K if self.opt_flags & F.OPT_INTRINSIC else 0)

    def desc(self) -> F.ProblemDesc:
        d = F.ProblemDesc()
        d.camera_model = self.camera_model
        d.opt_flags = self.opt_flags
        d.intr_dim = self.K
        d.intr_vars_present = 1 if self.intr_vars_present else 0
        d.ncam, d.npts, d.nobs = self.ncam, self.npts, self.nobs
        d.pose, d.intr, d.pt = F.ptr(self.pose), F.ptr(self.intr), F.ptr(self.pt)
        d.obs_cam, d.obs_pt, d.obs_uv = F.ptr(self.obs_cam), F.ptr(self.obs_pt), F.ptr(self.obs_uv)
        d.loss_id = self.loss_id
        lp = tuple(self.loss_params) + (0.0,) * (4 - len(self.loss_params))
        d.loss_params = (C.c_double * 4)(*lp[:4])
        d.pose_fixed = F.ptr(self.pose_fixed)
        d.intr_fixed = F.ptr(self.intr_fixed)
        d.pt_fixed = F.ptr(self.pt_fixed)
        if self.obs_loss is not None:
            tab = (F.LossSpec * len(self.loss_table))()
            for i, row in enumerate(self.loss_table):
                tab[i].loss_id = int(row[0])
                prm = tuple(row[1:]) + (0.0,) * 4
                tab[i].params = (C.c_double * 4)(*prm[:4])
            self._loss_table_c = tab   # keep alive while the descriptor is in use
            d.obs_loss = F.ptr(self.obs_loss)
            d.loss_table = C.cast(tab, C.c_void_p)
            d.n_losses = len(self.loss_table)
        return d


class Context:
    def __init__(self, lib, prefix: str, device: int = 0, rank: int = 0, nranks: int = 1, nccl_unique_id: Optional[bytes] = None):
        self._lib, self._p = lib, prefix
        desc = F.CtxDesc()
        desc.device, desc.rank, desc.nranks = device, rank, nranks
        self._uid = None
        if nccl_unique_id is not None:
            self._uid = C.create_string_buffer(bytes(nccl_unique_id), 128)
            desc.nccl_unique_id = C.cast(self._uid, C.c_void_p)
        h = C.c_void_p()
        st = self._fn("ctx_create")(C.byref(desc), C.byref(h))
        if st != F.OK:
            raise F.ApexError(st, "ctx_create")
        self._h = h
        self.problem: Optional[BAProblem] = None
        self.dims: Optional[F.Dims] = None

    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _check(self, st):
        if st != F.OK:
            msg = self._fn("last_error")(self._h)
            raise F.ApexError(st, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._fn("ctx_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- problem -------------------------------------------------------------------------------
    def upload(self, prob: BAProblem):
        d = prob.desc()
        self._check(self._fn("problem_upload")(self._h, C.byref(d)))
        self.problem = prob
        dims = F.Dims()
        self._check(self._fn("get_dims")(self._h, C.byref(dims)))
        self.dims = dims
        return self

    def params_upload(self, pose=None, intr=None, pt=None):
        pose, intr, pt = F.as_f64(pose), F.as_f64(intr), F.as_f64(pt)
        self._check(self._fn("params_upload")(self._h, F.ptr(pose), F.ptr(intr), F.ptr(pt)))

    def params_download(self):
        p = self.problem
        pose = np.empty((p.ncam, 7)); intr = np.empty((p.ncam, p.K)); pt = np.empty((p.npts, 3))
        self._check(self._fn("params_download")(self._h, F.ptr(pose), F.ptr(intr), F.ptr(pt)))
        return pose, intr, pt

    # -- stages --------------------------------------------------------------------------------
    def linearize(self, lam: float):
        self._check(self._fn("linearize")(self._h, float(lam)))

    def cost(self) -> float:
        c = C.c_double()
        self._check(self._fn("cost")(self._h, C.byref(c)))
        return c.value

    def get_linearization(self):
        p = self.problem
        r = np.empty((p.nobs, 2)); jc = np.empty((p.nobs, 2, p.dc)); jp = np.empty((p.nobs, 2, 3))
        self._check(self._fn("get_linearization")(self._h, F.ptr(r), F.ptr(jc), F.ptr(jp)))
        return r, jc, jp

    def get_blocks(self, *extra):
        p = self.problem
        hcc = np.empty((p.ncam, p.dc, p.dc)); gc = np.empty((p.ncam, p.dc))
        hpp = np.empty((p.npts, 3, 3)); gp = np.empty((p.npts, 3)); hinv = np.empty((p.npts, 3, 3))
        self._check(self._fn("get_blocks")(self._h, F.ptr(hcc), F.ptr(gc), F.ptr(hpp), F.ptr(gp), F.ptr(hinv), *extra))
        return hcc, gc, hpp, gp, hinv

    def schur_matvec(self, x):
        p = self.problem
        x = F.as_f64(x).reshape(p.ncam * p.dc)
        y = np.empty_like(x)
        self._check(self._fn("schur_matvec")(self._h, F.ptr(x), F.ptr(y)))
        return y

    def schur_matvec_bench(self, reps: int = 20, flush_l2: bool = True) -> float:
        ms = C.c_double()
        self._check(self._fn("schur_matvec_bench")(self._h, int(reps), 1 if flush_l2 else 0, C.byref(ms)))
        return ms.value

    def dense_cholesky_bench(self, n: int, reps: int = 3) -> float:
        """Average milliseconds of the dense FP64 (DMMA) Cholesky of a synthetic n x n SPD matrix (measurement aid)."""
        ms = C.c_double()
        self._check(self._fn("dense_cholesky_bench")(self._h, int(n), int(reps), C.byref(ms)))
        return ms.value

    def solve_augmented(self, variant: int, lam: float, precond: int = F.PRECOND_SCHUR_JACOBI, cg_max_iterations: int = 200,
                        cg_tolerance: float = 1e-6):
        p = self.problem
        sc = np.empty((p.ncam, p.dc)); sp = np.empty((p.npts, 3))
        g = C.c_double(); it = C.c_int32()
        self._check(self._fn("solve_augmented")(self._h, variant, precond, cg_max_iterations, cg_tolerance, float(lam), F.ptr(sc), F.ptr(sp),
                                                C.byref(g), C.byref(it)))
        return sc, sp, g.value, it.value

    def get_step(self):
        """(step_cam [ncam,dc], step_pt [npts,3]) of the last solve / the last LM iteration (parity read-back)."""
        p = self.problem
        sc = np.empty((p.ncam, p.dc)); sp = np.empty((p.npts, 3))
        self._check(self._fn("get_step")(self._h, F.ptr(sc), F.ptr(sp)))
        return sc, sp

    def default_config(self, for_bundle_adjustment: bool = True) -> F.LmConfig:
        cfg = F.LmConfig()
        self._fn("lm_config_for_bundle_adjustment" if for_bundle_adjustment else "lm_config_default")(C.byref(cfg))
        return cfg

    def lm_solve(self, cfg: F.LmConfig, trace_cap: int = 256):
        res = F.LmResult()
        tr = (F.IterTrace * trace_cap)()
        self._check(self._fn("lm_solve")(self._h, C.byref(cfg), C.byref(res), tr, trace_cap))
        n = min(res.iterations, trace_cap)
        return res, [tr[i] for i in range(n)]

    def add_observer(self, on_step=None, on_optimization_complete=None):
        """OptObserver (src/observers/mod.rs:201-330): `on_step(ctx, metrics)` once per LM iteration with the tuple the reference's
        notify_observers_generic passes (metrics: F.ObserverMetrics, valid during the call only; ctx: this context, whose
        params_download may be called from inside), `on_optimization_complete(ctx, iterations)` once when a solve ends."""
        step_cb = F.ON_STEP((lambda user, h, m: on_step(self, m.contents)) if on_step else 0)
        done_cb = F.ON_COMPLETE((lambda user, h, n: on_optimization_complete(self, n)) if on_optimization_complete else 0)
        obs = F.Observer(step_cb, done_cb, None)
        self._observers = getattr(self, "_observers", []) + [(step_cb, done_cb)]   # keep the thunks alive while they are registered
        self._check(self._fn("add_observer")(self._h, C.byref(obs)))
        return self

    def clear_observers(self):
        self._check(self._fn("clear_observers")(self._h))
        self._observers = []
        return self

    def kernel_launches(self) -> int:
        return int(self._fn("kernel_launches")(self._h))

    def profile_enable(self, on: bool = True):
        self._check(self._fn("profile_enable")(self._h, 1 if on else 0))

    def profile_read(self) -> F.Profile:
        p = F.Profile()
        self._check(self._fn("profile_read")(self._h, C.byref(p)))
        return p


def shard_info(obs_pt, npts: int, nranks: int, rank: int):
    """(block, npts_local, nobs_local): landmark p is owned by rank (p // block) % nranks (host-only entry point)."""
    lib = F.load_library()
    obs_pt = np.ascontiguousarray(obs_pt, dtype=np.uint32)
    blk, npl, n = C.c_uint32(), C.c_uint32(), C.c_uint64()
    st = lib.apex_shard_info(int(npts), int(obs_pt.shape[0]), F.ptr(obs_pt), int(nranks), int(rank), C.byref(blk), C.byref(npl), C.byref(n))
    if st != F.OK:
        raise F.ApexError(st, "shard_info")
    return blk.value, npl.value, n.value


def layout_stats(prob: BAProblem, nranks: int = 1, rank: int = 0) -> F.LayoutStats:
    """Host-only: build (and self-check) the observation layout rank `rank` of `nranks` would upload."""
    lib = F.load_library()
    d = prob.desc()
    out = F.LayoutStats()
    st = lib.apex_layout_stats_compute(C.byref(d), int(nranks), int(rank), C.byref(out))
    if st != F.OK:
        raise F.ApexError(st, "layout_stats")
    return out


class GpuContext(Context):
    """Context on the product library csrc/libapex_gpu.so (sm_100a kernels). No CPU fallback."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_unique_id: Optional[bytes] = None):
        super().__init__(F.load_library(), "apex_", device, rank, nranks, nccl_unique_id)

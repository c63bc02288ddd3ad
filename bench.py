#!/usr/bin/env python
"""bench.py — BA LM iterations/sec on the Venice-1778-shaped problem (BASELINE.json metric), 1..8 B200.

A "step" is ONE Levenberg–Marquardt iteration of the bundle-adjustment inner loop (linearize + block assembly +
implicit-Schur PCG solve + manifold update + trial cost + accept/reject bookkeeping) on the synthetic
Venice-1778-shaped input (1,778 cameras / 993,923 landmarks / ~5.3 M observations, BAL camera, self-calibration,
Huber(1.0), LevenbergMarquardtConfig::for_bundle_adjustment() with the matrix-free Schur solver). The convergence
tolerances are zeroed so that exactly K iterations run; everything else is the reference preset.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). `value` is timed on the device (CUDA events on the solver's own stream, max over
ranks) with the problem resident in HBM; `e2e` is the same metric through the C ABI from host buffers (upload of the
SoA problem + solve + download of the variables, wall clock). `--impl reference` times the CPU restatement of the
reference (oracle/, kind "port": the Rust crate cannot be built in this image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "BA LM iters/sec (Venice-1778 shape)"
UNIT = "LM iterations/s"
CPU_SAMPLE_SCALE = 1.0 / 16.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="venice1778")
    ap.add_argument("--scale", type=float, default=1.0, help="development only: shrink the workload")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="0 skips the cpu_baseline leg")
    return ap.parse_args()


def lm_config(ctx, iterations):
    """for_bundle_adjustment() preset, matrix-free Schur + Schur-Jacobi, exactly `iterations` LM iterations."""
    from apex_solver_b200 import _ffi as F
    cfg = ctx.default_config(True)
    cfg.schur_variant = F.SCHUR_IMPLICIT
    cfg.schur_preconditioner = F.PRECOND_SCHUR_JACOBI
    cfg.max_iterations = max(iterations - 1, 0)  # iterations = max_iterations + 1 (levenberg_marquardt.rs:1015)
    cfg.cost_tolerance = 0.0
    cfg.parameter_tolerance = 0.0
    cfg.gradient_tolerance = 0.0
    return cfg


def matvec_bytes(nobs, npts, ncam, dc):
    """Algorithmic bytes of one Schur-operator application."""
    return nobs * (8 * 2 * (dc + 3) + 8) + npts * 48 + ncam * (8 * dc * dc + 16 * dc)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_leg(shape, steps, warmup, full_nobs):
    """The reference's CPU path (restated oracle, OpenMP over observations like rayon par_iter; the Schur operator is
    single-threaded like apply_schur_operator_fast) on a bounded sample of the workload."""
    from apex_solver_b200 import synth
    from oracle_backend import OracleContext, oracle_lib
    prob = synth.make_shape(shape, scale=CPU_SAMPLE_SCALE)
    lib = oracle_lib()
    lib.oracle_set_num_threads(int(os.cpu_count() or 1))   # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    cores = int(lib.oracle_num_threads())
    ctx = OracleContext().upload(prob)
    pose0, intr0, pt0 = prob.pose.copy(), prob.intr.copy(), prob.pt.copy()
    if warmup > 0:
        ctx.lm_solve(lm_config(ctx, min(warmup, 1)))  # one LM iteration warms caches; more only burns minutes
    ctx.params_upload(pose0, intr0, pt0)
    t0 = time.perf_counter()
    res, trace = ctx.lm_solve(lm_config(ctx, steps))
    dt = time.perf_counter() - t0
    sample_rate = res.iterations / dt
    scaled = sample_rate * prob.nobs / full_nobs
    sample = (f"{shape} at 1/{round(1 / CPU_SAMPLE_SCALE)} scale ({prob.ncam} cams / {prob.npts} pts / {prob.nobs} obs), {res.iterations} LM iterations, "
              f"{res.linear_iterations} PCG iterations in {dt:.1f} s = {sample_rate:.3f} it/s on the sample; value = that x nobs_sample/nobs_full "
              f"(per-iteration work is linear in observations)")
    return {"value": scaled, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}, dt * 1e3 / max(res.iterations, 1)


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from apex_solver_b200 import _ffi as F, synth

    # full-size shape bookkeeping without generating it (the CPU leg needs nobs of the full workload)
    if a.impl == "reference":
        if rank != 0:
            return 0
        full = synth.make_shape(a.shape, scale=a.scale)
        cb, ms_step = cpu_reference_leg(a.shape, a.steps, a.warmup, full.nobs)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{a.shape} BAL-shaped synthetic, self-calibration, Huber(1.0), LM for_bundle_adjustment preset, implicit Schur PCG "
                                       f"(cg 200 / 1e-6, Schur-Jacobi)", "ncam": full.ncam, "npts": full.npts, "nobs": full.nobs},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from apex_solver_b200.context import GpuContext

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.set_device(local_rank)
    lib = F.load_library()
    uid = None
    if world > 1:
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (F.C.c_ubyte * 128)()
            st = lib.apex_nccl_unique_id(raw)
            if st != 0:
                raise F.ApexError(st, "apex_nccl_unique_id")
            buf = torch.tensor(list(raw), dtype=torch.uint8)
        dist.broadcast(buf, src=0)
        uid = bytes(buf.tolist())

    prob = synth.make_shape(a.shape, scale=a.scale)  # identical on every rank (seeded)
    dc = prob.dc
    ctx = GpuContext(device=local_rank, rank=rank, nranks=world, nccl_unique_id=uid)
    ctx.upload(prob)
    pose0, intr0, pt0 = prob.pose.copy(), prob.intr.copy(), prob.pt.copy()

    # ---- warm-up: W untimed LM iterations, then restore the initial variables ----
    if a.warmup > 0:
        ctx.lm_solve(lm_config(ctx, a.warmup))
    ctx.params_upload(pose0, intr0, pt0)

    # ---- timed: exactly K LM iterations, device time, max over ranks ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    n0 = ctx.kernel_launches()
    res, trace = ctx.lm_solve(lm_config(ctx, a.steps))
    timed = ctx.profile_read()          # lm_device_ms: CUDA events around the solve on the solver's stream
    launches = ctx.kernel_launches() - n0
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    dev_ms = torch.tensor([timed.lm_device_ms], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(dev_ms.item())
    steps_done = res.iterations
    value = steps_done / (dev_ms * 1e-3)

    # ---- roofline pass: the same K steps again with an event pair around every launch of the operator kernel (this
    # disables the CUDA-graph replay of the PCG batches, so it is kept out of the timed run above) ----
    ctx.params_upload(pose0, intr0, pt0)
    barrier()
    ctx.profile_enable(True)
    ctx.lm_solve(lm_config(ctx, a.steps))
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- e2e: the call a user makes, from host buffers: upload + solve + download, wall clock, max over ranks ----
    h2d = sum(x.nbytes for x in (prob.pose, prob.intr, prob.pt, prob.obs_cam, prob.obs_pt, prob.obs_uv))
    barrier()
    t0 = time.perf_counter()
    ctx.upload(prob)
    res_e, _ = ctx.lm_solve(lm_config(ctx, a.steps))
    out = ctx.params_download()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    d2h = sum(x.nbytes for x in out)
    e2e = {"value": res_e.iterations / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // max(res_e.iterations, 1),
           "d2h_bytes_per_step": d2h // max(res_e.iterations, 1), "wall_s": e2e_s,
           "note": "problem upload (H2D of the SoA factor graph + tile/segment structure build on the host) + solve + download of all variables"}

    # ---- roofline of the dominant kernel: the persistent Schur-operator kernel (PCG hot loop) ----
    dims = ctx.dims
    b_mv = matvec_bytes(int(dims.nobs_local), int(dims.npts_local), prob.ncam, dc)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    mv_ms = prof.matvec_ms / max(prof.matvec_launches, 1)
    achieved = b_mv / (mv_ms * 1e-3) / 1e9 if prof.matvec_launches else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "matvec_traffic.json")
    if os.path.exists(tpath) and world == 1 and a.scale == 1.0:
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "schur_chunk_kernel<9, MATVEC>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_mv, "launches": int(prof.matvec_launches), "avg_launch_ms": mv_ms,
                "share_of_step": prof.matvec_ms / prof.lm_device_ms if prof.lm_device_ms else None,
                "measured": "second pass of the same K steps with a CUDA-event pair around every launch (PCG batches not graph-replayed)"}

    if rank == 0:
        cb = None
        if world == 1 and a.cpu_baseline:
            cb, _ = cpu_reference_leg(a.shape, 2, 1, prob.nobs)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps_done, "warmup": a.warmup,
                "ms_per_step": dev_ms / max(steps_done, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{a.shape} BAL-shaped synthetic (seed 0xA9E50003), BAL camera, self-calibration, Huber(1.0), LM "
                                       f"for_bundle_adjustment preset (lambda0 1e-3), implicit Schur PCG (cg 200 / 1e-6, Schur-Jacobi), convergence tolerances zeroed "
                                       f"to run exactly K iterations",
                           "ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs, "cam_dof": prob.ncam * dc,
                           "parallelism": f"landmarks+observations sharded block-cyclically over {world} rank(s), camera blocks replicated; per PCG iteration: " + ("NVLink peer-memory all-reduce fused with p.Ap" if (int(ctx.dims.flags) & 1) else ("NCCL all-reduce" if world > 1 else "no exchange")),
                           "l2": "inputs larger than L2 (Jacobian planes 1.0 GB per operator application vs 126 MB L2), no explicit flush",
                           "pcg_iterations": int(res.linear_iterations), "final_cost": res.final_cost, "initial_cost": res.initial_cost,
                           "accepted_steps": int(res.successful_steps)},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline}
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py — BA LM iterations/sec on the Venice-1778-shaped problem (BASELINE.json metric), 1..8 B200.

A "step" is ONE Levenberg–Marquardt iteration of the bundle-adjustment inner loop (linearize + block assembly +
Schur solve + manifold update + trial cost + accept/reject bookkeeping) on a synthetic BASELINE.json-shaped input
(default: Venice-1778 shape, 1,778 cameras / 993,923 landmarks / ~5.3 M observations, BAL camera, self-calibration,
Huber(1.0), LevenbergMarquardtConfig::for_bundle_adjustment() with the matrix-free Schur solver, cg 200 / 1e-6). The
convergence tolerances are zeroed so that exactly K iterations run; everything else is the reference preset.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape S] [--variant V]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). `value` is timed on the device (CUDA events on the solver's own stream, max over
ranks) with the problem resident in HBM; `e2e` is the same metric through the C ABI from host buffers with a FRESH
context (allocation + host-side structure build + H2D of structure and values + solve + D2H of all variables, wall clock).
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the Rust crate cannot be built in
this image) on the host cores ON THE SAME CONFIG; see cpu_arm() for what exactly is measured.
Other BASELINE.json configs: --shape final13682 (C5), --shape kb2000|ds2000 --variant explicit (C4),
--shape ladybug49 --variant explicit (C1), --shape trafalgar257 (C2).
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

UNIT = "LM iterations/s"
SHAPE_LABEL = {"venice1778": "Venice-1778", "final13682": "Final-13682", "trafalgar257": "Trafalgar-257", "ladybug49": "Ladybug-49",
               "kb2000": "KB-2000 (C4)", "ds2000": "DS-2000 (C4)"}
DGEMM_FP64_TFLOPS = 35.5  # cuBLAS DGEMM 8192^3 measured on this pool's B200 in round 1 (profiles/r01_cholesky_vs_cusolver.json)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="venice1778")
    ap.add_argument("--variant", default="implicit", choices=["implicit", "explicit", "explicit_pcg"])
    ap.add_argument("--scale", type=float, default=1.0, help="development only: shrink the workload")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="0 skips the cpu_baseline leg of the GPU arm")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds of CPU work the reference arm may spend on its timed steps")
    ap.add_argument("--parity-vs-n1", type=int, default=1, help="N > 1: re-solve 3 iterations on one rank and report the difference")
    return ap.parse_args()


def lm_config(ctx, iterations, variant="implicit", timeout=0.0):
    """for_bundle_adjustment() preset with the given Schur variant, exactly `iterations` LM iterations."""
    from apex_solver_b200 import _ffi as F
    cfg = ctx.default_config(True)
    cfg.schur_variant = {"implicit": F.SCHUR_IMPLICIT, "explicit": F.SCHUR_EXPLICIT, "explicit_pcg": F.SCHUR_EXPLICIT_PCG}[variant]
    cfg.schur_preconditioner = F.PRECOND_SCHUR_JACOBI
    cfg.max_iterations = max(iterations - 1, 0)  # iterations = max_iterations + 1 (levenberg_marquardt.rs:1015)
    cfg.cost_tolerance = 0.0
    cfg.parameter_tolerance = 0.0
    cfg.gradient_tolerance = 0.0
    cfg.timeout_seconds = timeout
    return cfg


def matvec_bytes(nobs, npts, ncam, dc):
    """Algorithmic bytes of one Schur-operator application."""
    return nobs * (8 * 2 * (dc + 3) + 8) + npts * 48 + ncam * (8 * dc * dc + 16 * dc)


def workload_text(shape, variant, scale):
    solver = {"implicit": "implicit (matrix-free) Schur PCG (cg 200 / 1e-6, Schur-Jacobi)", "explicit": "explicit Schur + dense FP64 Cholesky",
              "explicit_pcg": "explicit Schur + scalar-Jacobi PCG (cg 200 / 1e-6)"}[variant]
    from apex_solver_b200 import synth
    model = {0: "BAL", 2: "Kannala-Brandt", 3: "double-sphere"}.get(synth.SHAPES[shape][3], str(synth.SHAPES[shape][3]))
    loss = "Huber(1.0)" if synth.SHAPES[shape][4][0] == 3 else "Cauchy(1.0)"
    sc = "" if scale == 1.0 else f" at scale {scale:g}"
    return (f"{shape}{sc} synthetic (seed 0xA9E5000{synth.SHAPES[shape][5]}), {model} camera, self-calibration, {loss}, LM for_bundle_adjustment "
            f"preset (lambda0 1e-3), {solver}, convergence tolerances zeroed to run exactly K iterations")


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


def cpu_arm(prob, shape, variant, steps, warmup, budget_s, faithful_leg):
    """The reference's CPU path (oracle/, the restatement of the crate: OpenMP over observations like rayon par_iter)
    on the FULL configuration `prob`, all host cores.

    value   = LM iterations / wall seconds of `steps` iterations (fewer when they do not fit `budget_s`: the LM loop's
              own timeout, levenberg_marquardt.rs:213-317 `timeout`, ends the run after the iteration that crosses it),
              with the landmark sweeps of the matrix-free solver (operator, reduced gradient, Schur-Jacobi build) spread
              over the cores. The reference runs those sweeps on ONE thread (apply_schur_operator_fast,
              implicit_schur.rs:163-251), so this is a FAVOURABLE stand-in for the crate.
    faithful = the same with single-threaded sweeps as in the reference, one LM iteration (when asked for and affordable).
    """
    from oracle_backend import OracleVariant, oracle_lib
    lib = oracle_lib()
    lib.oracle_set_num_threads(int(os.cpu_count() or 1))   # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    cores = int(lib.oracle_num_threads())
    ctx = OracleVariant(parallel=True).upload(prob)
    x0 = (prob.pose.copy(), prob.intr.copy(), prob.pt.copy())
    w_done = 0
    if warmup > 0:
        r, _ = ctx.lm_solve(lm_config(ctx, 1, variant))  # one LM iteration warms caches and page tables; more only burns minutes
        w_done = r.iterations
    ctx.params_upload(*x0)
    t0 = time.perf_counter()
    res, trace = ctx.lm_solve(lm_config(ctx, steps, variant, timeout=budget_s))
    dt = time.perf_counter() - t0
    rate = res.iterations / dt
    out = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": (f"FULL config ({prob.ncam} cams / {prob.npts} pts / {prob.nobs} obs), {res.iterations} of {steps} requested LM iterations "
                      f"({res.linear_iterations} PCG iterations) in {dt:.1f} s wall, {w_done} warm-up iteration(s); landmark sweeps of the matrix-free solver on "
                      f"{cores} threads (favourable: the reference runs them on one thread)"),
           "steps_done": int(res.iterations), "warmup_done": int(w_done), "wall_s": dt, "same_config": True,
           "final_cost": res.final_cost, "pcg_iterations": int(res.linear_iterations)}
    if faithful_leg:
        from oracle_backend import OracleContext
        fctx = OracleContext().upload(prob)
        t0 = time.perf_counter()
        fres, _ = fctx.lm_solve(lm_config(fctx, 1, variant))
        fdt = time.perf_counter() - t0
        out["faithful"] = {"value": fres.iterations / fdt, "unit": UNIT, "steps_done": int(fres.iterations), "wall_s": fdt,
                           "note": "landmark sweeps on ONE thread as in the reference (apply_schur_operator_fast); linearisation and cost on all cores"}
    return out, res


def cpu_scale_for(shape, variant):
    """Scale of the CPU arm's workload: full size wherever the oracle can run it in minutes. The explicit variants form a
    dense cam_dof^2 matrix in a sequential landmark loop (explicit_schur.rs:800-898): C4 (28 000^2) is run at 1/10."""
    if variant != "implicit" and shape in ("kb2000", "ds2000", "venice1778", "final13682"):
        return 0.1
    if shape == "final13682":
        return 0.125
    return 1.0


def parity_vs_one_rank(a, ctx, prob, x0, rank, local_rank, dist, barrier, iterates=3):
    """N > 1 correctness evidence inside the bench line. Free-running N-rank and 1-rank trajectories drift apart through
    rounding (truncated PCG on cond(S) ~1e10 amplifies another summation order to 1e-3 in the cost after a few LM iterations,
    also between two CPU runs of the reference algorithm), so the comparison is TEACHER-FORCED: at each of `iterates`
    iterates of the single-rank trajectory both contexts start from the same parameters, damping and nu and run ONE LM
    iteration of the bench configuration. Reported: accept flag / PCG iteration count equality, relative difference of the
    trial cost and of rho, and the conditioning-free BACKWARD agreement of the two camera steps
    ||S (dc_N - dc_1)|| / ||S dc_1|| (S applied by the single-rank context)."""
    import torch
    from apex_solver_b200.context import GpuContext
    c1 = GpuContext(device=local_rank).upload(prob) if rank == 0 else None
    x = [np.ascontiguousarray(v) for v in x0]
    lam, nu = 1e-3, 2.0
    rows = []
    for it in range(iterates):
        ctx.params_upload(*x)
        cfg = lm_config(ctx, 1, a.variant)
        cfg.damping, cfg.damping_nu = lam, nu
        barrier()
        rn, tn = ctx.lm_solve(cfg)
        dcn, _ = ctx.get_step()
        state = torch.zeros(2, dtype=torch.float64)
        if rank == 0:
            c1.params_upload(*x)
            cfg1 = lm_config(c1, 1, a.variant)
            cfg1.damping, cfg1.damping_nu = lam, nu
            r1, t1 = c1.lm_solve(cfg1)
            dc1, _ = c1.get_step()
            x_next = c1.params_download()
            c1.params_upload(*x)
            c1.linearize(lam)
            s_ref = c1.schur_matvec(dc1.ravel())
            s_diff = c1.schur_matvec((dcn - dc1).ravel())
            a_, b_ = tn[0], t1[0]
            rows.append({"iterate": it, "lambda": lam, "accept_n": int(a_.accepted), "accept_1": int(b_.accepted), "pcg_n": int(a_.ls_iter), "pcg_1": int(b_.ls_iter),
                         "rel_cost_at_iterate": abs(rn.initial_cost - r1.initial_cost) / abs(r1.initial_cost),
                         "rel_gradient_norm": abs(a_.gradient_norm - b_.gradient_norm) / abs(b_.gradient_norm),
                         "rel_trial_cost": abs(a_.new_cost - b_.new_cost) / abs(b_.new_cost), "abs_rho": abs(a_.tr_ratio - b_.tr_ratio),
                         "backward_step_agreement": float(np.linalg.norm(s_diff) / max(np.linalg.norm(s_ref), 1e-300))})
            x = [np.ascontiguousarray(v) for v in x_next]
            state = torch.tensor([r1.final_damping, r1.final_damping_nu], dtype=torch.float64)
        for k in range(3):   # rank 0's next iterate to everybody
            t = torch.from_numpy(x[k])
            dist.broadcast(t, src=0)
        dist.broadcast(state, src=0)
        lam, nu = float(state[0]), float(state[1])
    if rank != 0:
        return None
    c1.close()
    return {"method": "teacher-forced: N-rank and single-rank contexts run ONE LM iteration each from the same state, at 3 iterates of the single-rank trajectory",
            "iterates": rows,
            "accept_pattern_equal": all(r["accept_n"] == r["accept_1"] for r in rows),
            "pcg_iterations_equal": all(r["pcg_n"] == r["pcg_1"] for r in rows),
            "max_rel_cost_at_iterate": max(r["rel_cost_at_iterate"] for r in rows),
            "max_rel_trial_cost_diff": max(r["rel_trial_cost"] for r in rows),
            "max_backward_step_agreement": max(r["backward_step_agreement"] for r in rows)}


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from apex_solver_b200 import _ffi as F, synth
    metric = f"BA LM iters/sec ({SHAPE_LABEL.get(a.shape, a.shape)} shape)"

    if a.impl == "reference":
        if rank != 0:
            return 0
        cscale = a.scale * cpu_scale_for(a.shape, a.variant)
        prob = synth.make_shape(a.shape, scale=cscale)
        full_nobs = prob.nobs if cscale == a.scale else synth.make_shape(a.shape, scale=a.scale).nobs
        cb, res = cpu_arm(prob, a.shape, a.variant, a.steps, a.warmup, a.cpu_budget, faithful_leg=(a.variant == "implicit" and prob.nobs <= 8_000_000))
        value = cb["value"]
        if cscale != a.scale:   # the oracle cannot run this config at full size in minutes: say so, scale by observations
            cb["same_config"] = False
            cb["sample"] = f"scale {cscale:g} of the workload, value = sample rate x nobs_sample / nobs_full; " + cb["sample"]
            value = cb["value"] * prob.nobs / full_nobs
            cb["value"] = value
        line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": cb["steps_done"], "warmup": cb["warmup_done"],
                "steps_requested": a.steps, "warmup_requested": a.warmup,
                "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_text(a.shape, a.variant, a.scale), "ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs,
                           "cpu_budget_s": a.cpu_budget, "pcg_iterations": cb["pcg_iterations"], "final_cost": cb["final_cost"]},
                "cpu_baseline": cb,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    def make_uid():
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (F.C.c_ubyte * 128)()
            st = lib.apex_nccl_unique_id(raw)
            if st != 0:
                raise F.ApexError(st, "apex_nccl_unique_id")
            buf = torch.tensor(list(raw), dtype=torch.uint8)
        dist.broadcast(buf, src=0)
        return bytes(buf.tolist())

    prob = synth.make_shape(a.shape, scale=a.scale)  # identical on every rank (seeded)
    dc = prob.dc
    ctx = GpuContext(device=local_rank, rank=rank, nranks=world, nccl_unique_id=make_uid())
    ctx.upload(prob)
    x0 = (prob.pose.copy(), prob.intr.copy(), prob.pt.copy())

    # ---- warm-up: W untimed LM iterations, then restore the initial variables ----
    if a.warmup > 0:
        ctx.lm_solve(lm_config(ctx, a.warmup, a.variant))
    ctx.params_upload(*x0)

    # ---- timed: exactly K LM iterations, device time, max over ranks ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    n0 = ctx.kernel_launches()
    res, trace = ctx.lm_solve(lm_config(ctx, a.steps, a.variant))
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

    # ---- roofline pass: the same K steps again with an event pair around every launch of the dominant kernel (this
    # disables the CUDA-graph replay of the PCG batches, so it is kept out of the timed run above) ----
    ctx.params_upload(*x0)
    barrier()
    ctx.profile_enable(True)
    ctx.lm_solve(lm_config(ctx, a.steps, a.variant))
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- N > 1: teacher-forced parity of the N-rank solve against a single-rank solve (rank 0's own second context) ----
    parity = None
    if world > 1 and a.parity_vs_n1:
        parity = parity_vs_one_rank(a, ctx, prob, x0, rank, local_rank, dist, barrier)

    # ---- e2e: the call a user makes, from host buffers: upload (host-side structure build + H2D of structure and values) +
    # solve + D2H of all variables, wall clock, max over ranks. FIRST call on a context that has never seen a problem: its
    # page-locked staging arrays and device buffers are allocated inside the timed region (a "cold upload"; the context itself -
    # stream, NCCL communicator - exists, like the CUDA context of the process). warm_value: the same call again, buffers reused.
    def e2e_run(c):
        barrier()
        t0 = time.perf_counter()
        c.upload(prob)
        r, _ = c.lm_solve(lm_config(c, a.steps, a.variant))
        out = c.params_download()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return r, out, float(dt.item())

    t_ctx = time.perf_counter()
    ectx = GpuContext(device=local_rank, rank=rank, nranks=world, nccl_unique_id=make_uid())
    t_ctx = time.perf_counter() - t_ctx
    res_e, out, cold_s = e2e_run(ectx)
    up_bytes = int(ectx.profile_read().upload_h2d_bytes)
    res_w, _, warm_s = e2e_run(ectx)
    ectx.close()
    d2h = sum(x.nbytes for x in out)
    e2e = {"value": res_e.iterations / cold_s, "unit": UNIT, "h2d_bytes_per_step": up_bytes // max(res_e.iterations, 1),
           "d2h_bytes_per_step": d2h // max(res_e.iterations, 1), "wall_s": cold_s, "h2d_bytes_total": up_bytes,
           "raw_problem_bytes": sum(x.nbytes for x in (prob.pose, prob.intr, prob.pt, prob.obs_cam, prob.obs_pt, prob.obs_uv)),
           "warm_value": res_w.iterations / warm_s, "warm_wall_s": warm_s, "context_create_s": t_ctx,
           "note": "value: FIRST upload + solve + download on a new context (staging and device buffers allocated inside the timed region; context "
                   "creation - stream, NCCL communicator: context_create_s - outside, it is per-process setup); h2d bytes = structure + values actually "
                   "copied (this rank's shard); warm_value: the same call again on that context"}

    # ---- roofline of the dominant kernel ----
    dims = ctx.dims
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if a.variant == "explicit" and prof.cholesky_factorizations:
        n = int(prof.cholesky_n)
        ms = prof.cholesky_ms / prof.cholesky_factorizations
        flops = n ** 3 / 3.0
        ach = flops / (ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "dense FP64 Cholesky (chol_potrf/trsm/syrk128: DMMA.8x8x4)", "achieved": ach, "peak": DGEMM_FP64_TFLOPS,
                    "unit": "TFLOP/s", "frac": ach / DGEMM_FP64_TFLOPS, "traffic": None,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool in round 1 (MEASURED_PEAKS.json holds no FP64 figure)",
                    "algorithmic_flops_per_launch": flops, "n": n, "launches": int(prof.cholesky_factorizations), "avg_launch_ms": ms,
                    "share_of_step": prof.cholesky_ms / prof.lm_device_ms if prof.lm_device_ms else None,
                    "schur_form_ms": prof.schur_form_ms / max(prof.schur_forms, 1), "schur_form_share_of_step": prof.schur_form_ms / prof.lm_device_ms if prof.lm_device_ms else None}
    else:
        b_mv = matvec_bytes(int(dims.nobs_local), int(dims.npts_local), prob.ncam, dc)
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        mv_ms = prof.matvec_ms / max(prof.matvec_launches, 1)
        achieved = b_mv / (mv_ms * 1e-3) / 1e9 if prof.matvec_launches else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "matvec_traffic.json")
        if os.path.exists(tpath) and world == 1 and a.scale == 1.0 and a.shape == "venice1778":
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": "Schur operator kernel (MATVEC), dc = %d" % dc, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak if peak else None, "traffic": traffic,
                    "traffic_source": "profiles/matvec_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload (static file, not measured in this run)" if traffic else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": b_mv, "launches": int(prof.matvec_launches), "avg_launch_ms": mv_ms,
                    "share_of_step": prof.matvec_ms / prof.lm_device_ms if prof.lm_device_ms else None,
                    "measured": "second pass of the same K steps with a CUDA-event pair around every launch of the operator kernel (PCG batches not graph-replayed); "
                                "the second pass of the deterministic flush (the camera windows' partial rows, 2-3 % of the operator's bytes) runs inside the fused PCG tail kernel and is outside the pair"}

    if rank == 0:
        cb = None
        if world == 1 and a.cpu_baseline:
            cscale = a.scale * cpu_scale_for(a.shape, a.variant)
            cprob = prob if cscale == a.scale else synth.make_shape(a.shape, scale=cscale)
            cb, _ = cpu_arm(cprob, a.shape, a.variant, 2, 1, 60.0, faithful_leg=False)
            if cscale != a.scale:
                cb["same_config"] = False
                cb["sample"] = f"scale {cscale:g} of the workload, value = sample rate x nobs_sample / nobs_full; " + cb["sample"]
                cb["value"] = cb["value"] * cprob.nobs / prob.nobs
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps_done, "warmup": a.warmup,
                "ms_per_step": dev_ms / max(steps_done, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(a.shape, a.variant, a.scale),
                           "ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs, "cam_dof": prob.ncam * dc,
                           "parallelism": f"landmarks+observations sharded block-cyclically over {world} rank(s), camera blocks replicated; per PCG iteration: " + ("NVLink peer-memory all-reduce fused with p.Ap" if (int(ctx.dims.flags) & 1) else ("NCCL all-reduce" if world > 1 else "no exchange")),
                           "l2": "inputs larger than L2 (Jacobian planes 1.0 GB per operator application vs 126 MB L2), no explicit flush",
                           "pcg_iterations": int(res.linear_iterations), "final_cost": res.final_cost, "initial_cost": res.initial_cost,
                           "accepted_steps": int(res.successful_steps),
                           "note": "with the reference preset (cg 200 / 1e-6) PCG runs into its 200-iteration cap on almost every LM iteration of this workload, so LM it/s is "
                                   "essentially 1 / (200 operator applications + one linearisation)"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline}
        if parity is not None:
            line["config"]["parity_vs_n1"] = parity
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Development aid: run BASELINE.json configs (optionally scaled) through the GPU path and print timing / convergence."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
a = ap.parse_args()
RUNS = [  # name, shape, scale, variant, iterations
    ("C1 ladybug49 explicit", "ladybug49", 1.0, F.SCHUR_EXPLICIT, 20),
    ("C1 ladybug49 explicit_pcg (as dispatched)", "ladybug49", 1.0, F.SCHUR_EXPLICIT_PCG, 20),
    ("C2 trafalgar257 implicit", "trafalgar257", 1.0, F.SCHUR_IMPLICIT, 20),
    ("C2 trafalgar257 explicit", "trafalgar257", 1.0, F.SCHUR_EXPLICIT, 20),
    ("C4 kb2000 x0.1 explicit", "kb2000", 0.1, F.SCHUR_EXPLICIT, 5),
    ("C4 ds2000 x0.1 explicit", "ds2000", 0.1, F.SCHUR_EXPLICIT, 5),
    ("C3 venice1778 explicit", "venice1778", 1.0, F.SCHUR_EXPLICIT, 2),
    ("C4 kb2000 x0.5 explicit", "kb2000", 0.5, F.SCHUR_EXPLICIT, 2),
    ("C5 final13682 x0.5 implicit", "final13682", 0.5, F.SCHUR_IMPLICIT, 2),
]
for name, shape, scale, variant, iters in RUNS:
    if a.only and a.only not in name:
        continue
    try:
        t = time.time(); prob = synth.make_shape(shape, scale=scale); tg = time.time() - t
        t = time.time(); g = GpuContext().upload(prob); tu = time.time() - t
        cfg = g.default_config(True); cfg.schur_variant = variant; cfg.max_iterations = iters - 1
        cfg.cost_tolerance = 0.0; cfg.parameter_tolerance = 0.0; cfg.gradient_tolerance = 0.0
        g.profile_enable(True)
        t = time.time(); res, tr = g.lm_solve(cfg); dt = time.time() - t
        p = g.profile_read()
        print(json.dumps({"run": name, "ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs, "dc": prob.dc, "gen_s": round(tg, 2), "upload_s": round(tu, 2),
                          "iters": res.iterations, "lm_s": round(dt, 4), "it_per_s": round(res.iterations / dt, 3), "status": res.status,
                          "cost0": res.initial_cost, "cost1": res.final_cost, "accepted": res.successful_steps, "pcg": res.linear_iterations,
                          "matvec_ms_avg": round(p.matvec_ms / max(p.matvec_launches, 1), 4), "iter_ms": [round(x.iter_time_ms, 1) for x in tr][:6]}), flush=True)
        g.close()
    except Exception as e:
        print(json.dumps({"run": name, "error": repr(e)[:300]}), flush=True)

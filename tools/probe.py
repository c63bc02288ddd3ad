"""Quick GPU probe (development aid): upload a BASELINE.json shape, time linearize / Schur operator / LM."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="venice1778")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--variant", type=int, default=F.SCHUR_IMPLICIT)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--pts-frac", type=float, default=1.0, help="keep all cameras, this fraction of the landmarks (what one of 1/frac ranks holds)")
a = ap.parse_args()
t = time.time()
if a.pts_frac != 1.0:
    ncam, npts, track, model, loss, cfgi = synth.SHAPES[a.shape]
    prob = synth.make_problem(ncam, int(npts * a.pts_frac), track, camera_model=model, loss=loss, seed=0xA9E50000 + cfgi)
else:
    prob = synth.make_shape(a.shape, scale=a.scale)
tgen = time.time() - t
t = time.time(); g = GpuContext().upload(prob); tup = time.time() - t
dc = prob.dc
bytes_mv = prob.nobs * (8 * 2 * (dc + 3) + 8) + prob.npts * 48 + prob.ncam * (8 * dc * dc + 16 * dc)
out = {"shape": a.shape, "ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs, "gen_s": tgen, "upload_s": tup, "bytes_mv": bytes_mv}
t = time.time(); g.linearize(1e-3); out["linearize_first_ms"] = (time.time() - t) * 1e3
t = time.time(); g.linearize(1e-3); out["linearize_ms"] = (time.time() - t) * 1e3
t = time.time(); c = g.cost(); out["cost_ms"] = (time.time() - t) * 1e3; out["cost"] = c
for flush in (1, 0):
    ms = g.schur_matvec_bench(a.reps, bool(flush))
    out["matvec_ms_flush%d" % flush] = ms
    out["matvec_GBs_flush%d" % flush] = bytes_mv / ms / 1e6
cfg = g.default_config(True); cfg.schur_variant = a.variant; cfg.max_iterations = a.iters
n0 = g.kernel_launches()
t = time.time(); res, tr = g.lm_solve(cfg); dt = time.time() - t
out.update({"lm_s": dt, "lm_iters": res.iterations, "lm_it_per_s": res.iterations / dt, "status": res.status, "cost0": res.initial_cost,
            "cost1": res.final_cost, "pcg_iters": res.linear_iterations, "launches": g.kernel_launches() - n0,
            "trace": [(x.cost, x.accepted, x.ls_iter, round(x.iter_time_ms, 2)) for x in tr]})
print(json.dumps(out))

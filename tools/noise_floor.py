"""Development aid: per-iteration relative cost differences GPU-vs-oracle next to the oracle's own rounding floor
(oracle vs the same oracle accumulating in reverse order)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from apex_solver_b200 import _ffi as F, synth
from oracle_backend import OracleContext, oracle_lib

use_gpu = "--no-gpu" not in sys.argv
if use_gpu:
    from apex_solver_b200.context import GpuContext

def run(ctx, variant, max_it, cg_it=200, lam=1e-3):
    cfg = ctx.default_config(True); cfg.schur_variant = variant; cfg.max_iterations = max_it; cfg.cg_max_iterations = cg_it; cfg.damping = lam
    return ctx.lm_solve(cfg)

def rel(a, b): return abs(a - b) / abs(b)

cases = [("bal16 selfcal", dict(ncam=16, npts=600, mean_track=4.0, seed=7)),
         ("bal16 ba", dict(ncam=16, npts=600, mean_track=4.0, seed=7, self_calibration=False)),
         ("ladybug49", None)]
for name, kw in cases:
    prob = synth.make_shape("ladybug49") if kw is None else synth.make_problem(**kw)
    for variant in (F.SCHUR_EXPLICIT, F.SCHUR_IMPLICIT, F.SCHUR_EXPLICIT_PCG):
        L = oracle_lib()
        L.oracle_set_reverse_order(0); o = OracleContext().upload(prob); ro, to = run(o, variant, 8)
        L.oracle_set_reverse_order(1); o2 = OracleContext().upload(prob); r2, t2 = run(o2, variant, 8)
        L.oracle_set_reverse_order(0)
        line = f"{name:14s} v{variant} floor: " + " ".join("%.1e" % rel(a.cost, b.cost) for a, b in zip(t2, to))
        print(line, "| iters", r2.iterations, ro.iterations, "pcg", [x.ls_iter for x in t2][:4], [x.ls_iter for x in to][:4])
        if use_gpu:
            g = GpuContext().upload(prob); rg, tg = run(g, variant, 8)
            print(f"{name:14s} v{variant} gpu  : " + " ".join("%.1e" % rel(a.cost, b.cost) for a, b in zip(tg, to)), "| iters", rg.iterations, "pcg", [x.ls_iter for x in tg][:4],
                  "acc", [x.accepted for x in tg], [x.accepted for x in to])

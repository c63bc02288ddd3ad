"""Regression fixtures (NOT reference outputs: the Rust reference cannot be built in this image).

Runs the test oracle on small seeded problems - one per camera model, both Schur variants - and stores what the LM loop
produces (per-iteration cost, accept pattern, PCG iterations, status, final parameters digest) together with unit-stage
values (cost, gradient norm, one Schur operator application). `tests/test_golden.py` checks that the oracle still
reproduces them (CPU) and that the GPU path matches them (B200). Regenerate with
    python tests/golden/make_golden.py
after an intended change of the algorithm, and say so in the commit message."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from apex_solver_b200 import _ffi as F, synth  # noqa: E402
from oracle_backend import OracleContext  # noqa: E402

CASES = [
    ("bal_selfcal_implicit", dict(camera_model=F.CAM_BAL, self_calibration=True, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_IMPLICIT),
    ("bal_ba_explicit", dict(camera_model=F.CAM_BAL, self_calibration=False, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_EXPLICIT),
    ("pinhole_selfcal_explicit", dict(camera_model=F.CAM_PINHOLE, self_calibration=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_EXPLICIT),
    ("kb_selfcal_explicit", dict(camera_model=F.CAM_KANNALA_BRANDT, self_calibration=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_EXPLICIT),
    ("ds_ba_implicit", dict(camera_model=F.CAM_DOUBLE_SPHERE, self_calibration=False, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_IMPLICIT),
    ("radtan_selfcal_explicit", dict(camera_model=F.CAM_RADTAN, self_calibration=True, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_EXPLICIT),
    ("ucm_ba_implicit", dict(camera_model=F.CAM_UCM, self_calibration=False, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_IMPLICIT),
    ("eucm_selfcal_explicit", dict(camera_model=F.CAM_EUCM, self_calibration=True, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_EXPLICIT),
    ("fov_ba_explicit", dict(camera_model=F.CAM_FOV, self_calibration=False, loss=(F.LOSS_L2,)), F.SCHUR_EXPLICIT),
    ("ftheta_ba_implicit", dict(camera_model=F.CAM_FTHETA, self_calibration=False, loss=(F.LOSS_HUBER, 1.0)), F.SCHUR_IMPLICIT),
]
SHAPE = dict(ncam=8, npts=120, mean_track=4.0)


def problem(kw, seed):
    return synth.make_problem(SHAPE["ncam"], SHAPE["npts"], SHAPE["mean_track"], seed=seed, **kw)


def run(ctx, variant, max_it=6):
    cfg = ctx.default_config(True)
    cfg.schur_variant = variant
    cfg.max_iterations = max_it
    cfg.cg_max_iterations = 300
    cfg.cg_tolerance = 1e-10
    return ctx.lm_solve(cfg)


def record(ctx, prob, variant):
    n = prob.ncam * prob.dc
    x = np.sin(0.37 * np.arange(1, n + 1))          # fixed, seed-free operator input
    out = {"cost0": ctx.cost()}
    ctx.linearize(1e-3)
    y = ctx.schur_matvec(x)
    out["matvec_norm"] = float(np.linalg.norm(y))
    out["matvec_head"] = [float(v) for v in y[:6]]
    res, tr = run(ctx, variant)
    out.update(status=int(res.status), iterations=int(res.iterations), initial_cost=res.initial_cost, final_cost=res.final_cost,
               costs=[t.cost for t in tr], accepted=[int(t.accepted) for t in tr], pcg=[int(t.ls_iter) for t in tr])
    pose, intr, pt = ctx.params_download()
    out["param_norms"] = [float(np.linalg.norm(pose)), float(np.linalg.norm(intr)), float(np.linalg.norm(pt))]
    return out


def main():
    golden = {"shape": SHAPE, "cases": {}}
    for i, (name, kw, variant) in enumerate(CASES):
        prob = problem(kw, 100 + i)
        golden["cases"][name] = record(OracleContext().upload(prob), prob, variant)
        print(name, golden["cases"][name]["iterations"], golden["cases"][name]["final_cost"])
    with open(os.path.join(HERE, "oracle_lm_small.json"), "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()

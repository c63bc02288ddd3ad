"""Parity of the sm_100a path (through the C ABI of include/apex_gpu.h) against the TEST ORACLE on the same
seeded inputs. `-m gpu` only. Tolerances follow BASELINE.json north_star: per-iteration cost <= 1e-9
relative, final cost / parameters <= 1e-6 relative, identical LM iteration count and termination status;
unit stages are held much tighter (element-wise FP64 work is compiled without FMA contraction, so it
differs from the CPU only through libm and summation order).
"""
import numpy as np
import pytest

from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
from oracle_backend import OracleContext, OracleVariant, oracle_lib
from parity_helpers import teacher_forced_parity

pytestmark = pytest.mark.gpu

COST_ITER_RTOL = 1e-9   # north_star: per-iteration cost
FINAL_RTOL = 1e-6       # north_star: final cost and parameters
# The reduced camera system of a self-calibrating BA problem is ill conditioned (cond ~1e10+, one gauge direction
# held only by lambda), so the reference algorithm itself is reproducible only down to a rounding floor: the
# oracle run twice with its block sums accumulated in opposite order differs by 1e-9..1e-5 in the per-iteration cost
# (tools/noise_floor.py). Where that measured floor is above 1e-9 the GPU is held to FLOOR_FACTOR x floor instead.
FLOOR_FACTOR = 10.0
# ... but never looser than this per-iteration cost tolerance: a case whose floor is worse tests nothing and gets a
# converged-PCG configuration instead (three cases ran at an effective 2e-3..2e-2).
TOL_CAP = 1e-5


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    return float(np.abs(a - b).max(initial=0.0) / scale)


def small_problem(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_HUBER, 1.0), ncam=12, npts=400, track=4.0, seed=7, **kw):
    return synth.make_problem(ncam, npts, track, camera_model=model, loss=loss, seed=seed, self_calibration=self_cal, **kw)


def pair(prob):
    return GpuContext().upload(prob), OracleContext().upload(prob)


def assert_blocks_close(g, o, prob, lam, tol=1e-11):
    hg, ho = g.get_blocks(), o.get_blocks()
    for a, b, what in zip(hg[:4], ho[:4], ("H_cc", "g_c", "H_pp", "g_p")):
        assert relerr(a, b) < tol, what
    # the guarded 3x3 inverse inherits the conditioning of its block: compare per block, scaled by cond
    damped = ho[2] + lam * np.eye(3)[None]
    cond = np.linalg.cond(damped)
    err = np.abs(hg[4] - ho[4]).max(axis=(1, 2)) / np.abs(ho[4]).max(axis=(1, 2))
    assert (err <= 1e-13 * np.maximum(cond, 1.0) + 1e-13).all(), "H_pp^-1"
    return cond


CASES = [
    ("bal_selfcal_huber", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_HUBER, 1.0))),
    ("bal_ba_huber", dict(model=F.CAM_BAL, self_cal=False, loss=(F.LOSS_HUBER, 1.0))),
    ("bal_selfcal_l2", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_L2,))),
    ("bal_selfcal_noloss", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_NONE,))),
    ("pinhole_selfcal_cauchy", dict(model=F.CAM_PINHOLE, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0))),
    ("kb_selfcal_cauchy", dict(model=F.CAM_KANNALA_BRANDT, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0))),
    ("ds_selfcal_cauchy", dict(model=F.CAM_DOUBLE_SPHERE, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0))),
    ("kb_ba_huber", dict(model=F.CAM_KANNALA_BRANDT, self_cal=False, loss=(F.LOSS_HUBER, 2.0))),
    ("radtan_selfcal_huber", dict(model=F.CAM_RADTAN, self_cal=True, loss=(F.LOSS_HUBER, 1.0))),      # dc = 15
    ("ucm_selfcal_cauchy", dict(model=F.CAM_UCM, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0))),          # dc = 11
    ("eucm_selfcal_huber", dict(model=F.CAM_EUCM, self_cal=True, loss=(F.LOSS_HUBER, 1.0))),          # dc = 12
    ("fov_selfcal_huber", dict(model=F.CAM_FOV, self_cal=True, loss=(F.LOSS_HUBER, 1.0))),            # dc = 11
    ("ftheta_selfcal_cauchy", dict(model=F.CAM_FTHETA, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0))),    # dc = 12
    ("radtan_ba_huber", dict(model=F.CAM_RADTAN, self_cal=False, loss=(F.LOSS_HUBER, 1.0))),
    ("bal_selfcal_andrews", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_ANDREWS, 3.0))),  # rho'' > 0: corrector 2nd branch
    ("bal_selfcal_tukey", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_TUKEY, 4.0))),
    ("bal_shuffled", dict(model=F.CAM_BAL, self_cal=True, shuffle_obs=True)),
    # the remaining loss functions (src/core/loss_functions.rs:236,585,674,759,1037,1132,1207,1316,1445), parameters as in their doc examples
    ("bal_selfcal_l1", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_L1,))),
    ("bal_selfcal_fair", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_FAIR, 1.3998))),
    ("bal_selfcal_geman_mcclure", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_GEMAN_MCCLURE, 1.0))),
    ("bal_selfcal_welsch", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_WELSCH, 2.9846))),
    ("bal_selfcal_ramsay_ea", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_RAMSAY_EA, 0.3))),
    ("bal_selfcal_trimmed_mean", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_TRIMMED_MEAN, 2.0))),
    ("bal_selfcal_lp_norm", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_LP_NORM, 1.5))),
    ("bal_selfcal_barron", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_BARRON, 1.0, 1.0))),
    ("kb_selfcal_barron_negative_alpha", dict(model=F.CAM_KANNALA_BRANDT, self_cal=True, loss=(F.LOSS_BARRON, -2.0, 1.5))),
    ("bal_selfcal_t_distribution", dict(model=F.CAM_BAL, self_cal=True, loss=(F.LOSS_T_DISTRIBUTION, 5.0))),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_linearize_blocks_cost_matvec(name, kw):
    prob = small_problem(**kw)
    g, o = pair(prob)
    lam = 1e-3
    assert abs(g.cost() - o.cost()) <= 1e-13 * abs(o.cost())
    g.linearize(lam); o.linearize(lam)
    rg, jcg, jpg = g.get_linearization()
    ro, jco, jpo = o.get_linearization()
    assert relerr(rg, ro) < 1e-13, "residuals"
    assert relerr(jcg, jco) < 1e-12, "camera Jacobian blocks"
    assert relerr(jpg, jpo) < 1e-12, "landmark Jacobian blocks"
    cond = assert_blocks_close(g, o, prob, lam)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(prob.ncam * prob.dc)
    yg, yo = g.schur_matvec(x), o.schur_matvec(x)
    assert relerr(yg, yo) < 1e-12 * max(cond.max(), 10.0), "Schur operator"
    assert relerr(g.schur_matvec(x), yg) < 1e-13  # default path: FP64 reductions into L2, order may differ run to run


@pytest.mark.parametrize("model", [F.CAM_BAL, F.CAM_KANNALA_BRANDT], ids=["bal", "kb"])
def test_per_block_loss_functions(model):
    """apex_problem_desc::obs_loss / loss_table (every ResidualBlock owns its own loss, src/core/residual_block.rs:97-123): four
    LossFunction instances drawn per observation. K1 (slot order), the cost kernel and the camera-major accumulations (which
    re-evaluate the corrector) must all pick the block's own loss: residuals, Jacobians, blocks and an LM run against the oracle."""
    import dataclasses
    # (no gross outliers and no redescending loss in the table: an un-weighted 50-pixel residual next to Tukey-zeroed blocks
    # leaves landmark blocks held by lambda alone, and two correct implementations then agree only to 1e-8 in the back-substitution)
    base = small_problem(model=model, ncam=14, npts=500, outlier_frac=0.0)
    table = [(F.LOSS_HUBER, 1.0), (F.LOSS_CAUCHY, 2.0), (F.LOSS_NONE,), (F.LOSS_FAIR, 1.5)]
    idx = np.random.default_rng(3).integers(0, len(table), base.nobs).astype(np.uint8)
    prob = dataclasses.replace(base, obs_loss=idx, loss_table=table, meta={})
    g, o = pair(prob)
    assert abs(g.cost() - o.cost()) <= 1e-13 * abs(o.cost())
    assert abs(g.cost() - GpuContext().upload(base).cost()) > 1e-6 * abs(o.cost()), "the table must matter"
    g.linearize(1e-3); o.linearize(1e-3)
    for a, b, what in zip(g.get_linearization(), o.get_linearization(), ("residuals", "camera Jacobians", "landmark Jacobians")):
        assert relerr(a, b) < 1e-12, what
    assert_blocks_close(g, o, prob, 1e-3)
    teacher_forced(prob, F.SCHUR_EXPLICIT, n_it=4, backward_tol=1e-10)
    teacher_forced(prob, F.SCHUR_IMPLICIT, n_it=4)


@pytest.mark.parametrize("model", [F.CAM_KANNALA_BRANDT, F.CAM_DOUBLE_SPHERE], ids=["kannala_brandt", "double_sphere"])
def test_shared_intrinsics_calibration_graph(model):
    """The reference's multi-observation / shared-intrinsics graphs (one ProjectionFactor with all its observations per camera over
    [pose_k, landmarks, intrinsics], src/factors/projection_factor.rs:184-364, tests/camera_*_integration.rs) on the reference's own
    inputs: the GPU path solves the per-camera reduced system projected onto the shared variable, the oracle the full normal
    equations by dense Cholesky like the reference's default solver. One teacher-forced iteration per iterate (step, cost, rho,
    damping), then the free-running LM against the oracle's and against the reference tests' recovery criteria."""
    from test_host_cpu import reference_calibration_problem
    from parity_helpers import one_iteration
    prob = reference_calibration_problem(model)
    g, o = pair(prob)
    lam, nu = 1e-3, 2.0
    for it in range(5):   # teacher-forced: both start every iteration from the oracle's iterate
        g.params_upload(*o.params_download())
        (ra, a, _), (rb, b, _) = one_iteration(g, F.SCHUR_EXPLICIT, lam, nu, 0, 0.0), one_iteration(o, F.SCHUR_EXPLICIT, lam, nu, 0, 0.0)
        (dca, dpa), (dcb, dpb) = g.get_step(), o.get_step()
        assert abs(ra.initial_cost - rb.initial_cost) <= 1e-13 * rb.initial_cost and abs(a.gradient_norm - b.gradient_norm) <= 1e-11 * b.gradient_norm, it
        # (two different algorithms on a planar-target calibration problem: agreement of the steps to 1e-8 / 2e-7 measured)
        assert relerr(dca, dcb) < 1e-5 and relerr(dpa, dpb) < 1e-5, (it, relerr(dca, dcb), relerr(dpa, dpb))
        assert np.array_equal(dca[:, 6:], np.tile(dca[0, 6:], (prob.ncam, 1))), "every camera's copy takes the shared step"
        assert abs(a.step_norm - b.step_norm) <= 1e-7 * b.step_norm and abs(a.predicted_reduction - b.predicted_reduction) <= 1e-7 * abs(b.predicted_reduction)
        assert a.accepted == b.accepted and abs(a.new_cost - b.new_cost) <= 1e-7 * b.new_cost and abs(ra.final_damping - rb.final_damping) <= 1e-6 * rb.final_damping, it
        lam, nu = rb.final_damping, rb.final_damping_nu
    g, o = pair(prob)
    kw = dict(max_it=100, cost_tolerance=1e-8, parameter_tolerance=1e-8, gradient_tolerance=1e-10, damping=1e-3)
    (rg, tg), (ro, to) = run_lm(g, F.SCHUR_EXPLICIT, **kw), run_lm(o, F.SCHUR_EXPLICIT, **kw)
    assert (rg.status, rg.iterations, [t.accepted for t in tg]) == (ro.status, ro.iterations, [t.accepted for t in to])
    assert abs(rg.final_cost - ro.final_cost) <= 1e-6 * max(ro.final_cost, 1e-9) + 1e-12
    (pg, ig, xg), (po, io, xo) = g.params_download(), o.params_download()
    assert np.array_equal(ig, np.tile(ig[0], (prob.ncam, 1))) and relerr(ig[0], io[0]) < 1e-6 and relerr(pg, po) < 1e-6 and relerr(xg, xo) < 1e-6
    truth = prob.meta["truth_intr"]
    assert rg.status in (0, 2, 3, 4) and (rg.initial_cost - rg.final_cost) / rg.initial_cost > 0.85 and np.sqrt(rg.final_cost / prob.nobs) < 2.0
    assert all(abs(ig[0, i] - truth[i]) / max(abs(truth[i]), 0.1) < 0.25 for i in range(4))
    with pytest.raises(F.ApexError) as e:   # the matrix-free solver has no shared-variable form
        run_lm(g, F.SCHUR_IMPLICIT, max_it=1)
    assert e.value.status == F.ERR_UNSUPPORTED


@pytest.mark.parametrize("variant", [F.SCHUR_EXPLICIT, F.SCHUR_IMPLICIT, F.SCHUR_EXPLICIT_PCG], ids=["explicit", "implicit", "explicit_pcg"])
@pytest.mark.parametrize("self_cal", [True, False], ids=["selfcal", "ba"])
def test_solve_augmented(variant, self_cal):
    prob = small_problem(self_cal=self_cal, ncam=10, npts=300)
    g, o = pair(prob)
    lam = 1e-2
    sg = g.solve_augmented(variant, lam, cg_max_iterations=500, cg_tolerance=1e-12)
    so = o.solve_augmented(variant, lam, cg_max_iterations=500, cg_tolerance=1e-12)
    assert abs(sg[2] - so[2]) <= 1e-12 * so[2], "gradient norm"
    # both solve the same SPD system to 1e-12: steps agree to solver accuracy
    assert relerr(sg[0], so[0]) < 1e-7, "camera step"
    assert relerr(sg[1], so[1]) < 1e-7, "landmark step"
    if variant == F.SCHUR_EXPLICIT:
        assert relerr(sg[0], so[0]) < 1e-9


@pytest.mark.parametrize("precond", [F.PRECOND_NONE, F.PRECOND_BLOCK_DIAGONAL, F.PRECOND_SCHUR_JACOBI], ids=["none", "blockdiag", "schurjacobi"])
def test_pcg_iteration_counts_match(precond):
    prob = small_problem(ncam=10, npts=300)
    g, o = pair(prob)
    sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, precond=precond, cg_max_iterations=30, cg_tolerance=1e-6)
    so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, precond=precond, cg_max_iterations=30, cg_tolerance=1e-6)
    assert sg[3] == so[3], "PCG iterations"
    # 30 unconverged CG iterations on an ill-conditioned system amplify rounding; the reordered oracle is the yardstick
    oracle_lib().oracle_set_reverse_order(1)
    try:
        s2 = OracleContext().upload(prob).solve_augmented(F.SCHUR_IMPLICIT, 1e-2, precond=precond, cg_max_iterations=30, cg_tolerance=1e-6)
    finally:
        oracle_lib().oracle_set_reverse_order(0)
    floor = max(relerr(s2[0], so[0]), relerr(s2[1], so[1]))
    assert relerr(sg[0], so[0]) <= max(1e-9, FLOOR_FACTOR * floor)
    assert relerr(sg[1], so[1]) <= max(1e-9, FLOOR_FACTOR * floor)


def run_lm(ctx, variant, max_it=8, cg_it=200, **cfgkw):
    cfg = ctx.default_config(True)
    cfg.schur_variant = variant
    cfg.max_iterations = max_it
    cfg.cg_max_iterations = cg_it
    for k, v in cfgkw.items():
        setattr(cfg, k, v)
    return ctx.lm_solve(cfg)


def oracle_lm_with_floor(prob, variant, **kw):
    """Oracle LM trajectory + its own rounding floor: the same oracle with its block sums accumulated in reverse order and
    with its landmark sweeps split over 3 threads (two reorderings; the floor is the larger deviation)."""
    ro, to = run_lm(OracleContext().upload(prob), variant, **kw)
    floor = [0.0] * len(to)
    for twin in (OracleVariant(reverse=True), OracleVariant(parallel=True, threads=3)):
        r2, t2 = run_lm(twin.upload(prob), variant, **kw)
        # the reference algorithm itself must take one path regardless of summation order, or the case is useless
        assert (r2.status, r2.iterations) == (ro.status, ro.iterations) and all(a.accepted == b.accepted for a, b in zip(t2, to)), \
            "the oracle's own trajectory depends on its summation order: pick another case"
        floor = [max(f, abs(a.cost - b.cost) / abs(b.cost)) for f, a, b in zip(floor, t2, to)]
    return ro, to, floor


def assert_lm_parity(rg, tg, ro, to, floor, strict=False):
    assert abs(rg.initial_cost - ro.initial_cost) <= 1e-13 * ro.initial_cost
    run_floor = max(floor)  # rounding noise random-walks along the trajectory: the yardstick is its maximum
    tol = COST_ITER_RTOL if strict else max(COST_ITER_RTOL, FLOOR_FACTOR * run_floor)
    assert tol <= TOL_CAP, f"rounding floor {run_floor:.1e}: this case tests nothing, give it a converged linear solve"
    if strict:
        assert run_floor < COST_ITER_RTOL, "case is meant to be well conditioned"
    assert rg.status == ro.status, "termination status"
    assert rg.iterations == ro.iterations, "LM iteration count"
    assert rg.successful_steps == ro.successful_steps and rg.unsuccessful_steps == ro.unsuccessful_steps
    assert rg.cost_evaluations == ro.cost_evaluations and rg.jacobian_evaluations == ro.jacobian_evaluations
    for a, b in zip(tg, to):
        assert a.accepted == b.accepted, f"accept/reject at iteration {b.iteration}"
        assert abs(a.cost - b.cost) <= tol * abs(b.cost), f"cost at iteration {b.iteration}: {a.cost} vs {b.cost} (floor {run_floor:.1e})"
    assert abs(rg.final_cost - ro.final_cost) <= max(FINAL_RTOL, FLOOR_FACTOR * run_floor) * abs(ro.final_cost)
    return tol


# PCG run to convergence: a truncated PCG (200 iterations at its cap) on an ill-conditioned reduced system amplifies rounding to
# 1e-3..1e-2 in the cost; converged, the same cases reproduce to 1e-9..1e-8
CONVERGED = dict(cg_it=3000, cg_tolerance=1e-12)
LM_CASES = [
    ("bal_selfcal_explicit", dict(model=F.CAM_BAL, self_cal=True), F.SCHUR_EXPLICIT, {}),
    # the headline mode with the reference preset (cg 200 / 1e-6): 4 LM iterations - from the 5th on the oracle's own twins differ by 1.6e-6
    ("bal_selfcal_implicit", dict(model=F.CAM_BAL, self_cal=True), F.SCHUR_IMPLICIT, dict(max_it=3)),
    ("bal_selfcal_explicit_pcg", dict(model=F.CAM_BAL, self_cal=True), F.SCHUR_EXPLICIT_PCG, CONVERGED),
    ("bal_ba_implicit_strict", dict(model=F.CAM_BAL, self_cal=False), F.SCHUR_IMPLICIT, {}),
    ("bal_ba_explicit", dict(model=F.CAM_BAL, self_cal=False), F.SCHUR_EXPLICIT, {}),
    ("kb_selfcal_explicit", dict(model=F.CAM_KANNALA_BRANDT, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_EXPLICIT, {}),
    ("ds_selfcal_explicit", dict(model=F.CAM_DOUBLE_SPHERE, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_EXPLICIT, {}),
    ("ds_selfcal_implicit", dict(model=F.CAM_DOUBLE_SPHERE, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_IMPLICIT, CONVERGED),
    ("pinhole_selfcal_implicit", dict(model=F.CAM_PINHOLE, self_cal=True), F.SCHUR_IMPLICIT, CONVERGED),
    ("radtan_selfcal_explicit", dict(model=F.CAM_RADTAN, self_cal=True), F.SCHUR_EXPLICIT, {}),
    ("ucm_selfcal_implicit", dict(model=F.CAM_UCM, self_cal=True, loss=(F.LOSS_CAUCHY, 1.0)), F.SCHUR_IMPLICIT, CONVERGED),
    ("eucm_ba_implicit", dict(model=F.CAM_EUCM, self_cal=False), F.SCHUR_IMPLICIT, {}),
    ("fov_selfcal_explicit", dict(model=F.CAM_FOV, self_cal=True), F.SCHUR_EXPLICIT, {}),
    ("ftheta_selfcal_explicit", dict(model=F.CAM_FTHETA, self_cal=True), F.SCHUR_EXPLICIT, {}),
]


@pytest.mark.parametrize("name,kw,variant,cg", LM_CASES, ids=[c[0] for c in LM_CASES])
def test_lm_solve_parity(name, kw, variant, cg):
    """Free-running LM: status, iteration count, accept pattern, evaluation counts identical; per-iteration cost within
    max(1e-9, 10 x the oracle's own rounding floor), which is asserted to stay below TOL_CAP."""
    prob = small_problem(ncam=16, npts=600, **kw)
    g = GpuContext().upload(prob)
    rg, tg = run_lm(g, variant, **cg)
    ro, to, floor = oracle_lm_with_floor(prob, variant, **cg)
    tol = assert_lm_parity(rg, tg, ro, to, floor, strict=name.endswith("_strict"))
    o = OracleContext().upload(prob)
    run_lm(o, variant, **cg)
    ptol = max(FINAL_RTOL, tol * 1e3)  # parameters move ~sqrt faster than the cost along flat directions
    for a, b, what in zip(g.params_download(), o.params_download(), ("poses", "intrinsics", "landmarks")):
        assert relerr(a, b) < ptol, what


def teacher_forced(prob, variant, n_it, fast_oracle=False, **kw):
    """tests/parity_helpers.py on the GPU: every link of one LM iteration against the oracle's arithmetic at each iterate
    of the oracle's trajectory. fast_oracle: the oracle's landmark sweeps run on all host cores (full-size cases)."""
    mk = (lambda: OracleVariant(parallel=True)) if fast_oracle else OracleContext
    rows = teacher_forced_parity(prob, variant, GpuContext().upload(prob), mk().upload(prob), mk().upload(prob), oracle_lib(),
                                 twin=OracleVariant(reverse=True, parallel=fast_oracle, threads=5 if fast_oracle else None).upload(prob), n_it=n_it, log=print, **kw)
    assert len(rows) == n_it
    return rows


@pytest.mark.parametrize("name,kw,variant,cg", LM_CASES, ids=[c[0] for c in LM_CASES])
def test_lm_teacher_forced(name, kw, variant, cg):
    teacher_forced(small_problem(ncam=16, npts=600, **kw), variant, n_it=6)


def test_lm_teacher_forced_ladybug49_full():
    """BASELINE.json configs[0] at full size (C1: 49 cams / 7 776 pts / 31.8k obs), explicit Schur + Cholesky."""
    rows = teacher_forced(synth.make_shape("ladybug49"), F.SCHUR_EXPLICIT, n_it=8)
    assert all(r["same_accept"] for r in rows)


def test_lm_teacher_forced_trafalgar257_full():
    """configs[1] at full size (C2: 257 cams / 65k pts / 226k obs), matrix-free PCG run to convergence."""
    teacher_forced(synth.make_shape("trafalgar257"), F.SCHUR_IMPLICIT, n_it=4, fast_oracle=True, cg_it=10000)


def test_lm_teacher_forced_venice_sixteenth():
    """configs[2] (C3, the bench workload) at 1/16 scale: 111 cams / 62k pts / 273k obs."""
    teacher_forced(synth.make_shape("venice1778", scale=1.0 / 16.0), F.SCHUR_IMPLICIT, n_it=3, fast_oracle=True, cg_it=10000)


def test_lm_teacher_forced_final_sixtyfourth():
    """configs[4] (C5, Final-13682 shape) at 1/64 scale, matrix-free PCG."""
    teacher_forced(synth.make_shape("final13682", scale=1.0 / 64.0), F.SCHUR_IMPLICIT, n_it=2, fast_oracle=True, cg_it=10000)


@pytest.mark.parametrize("shape", ["kb2000", "ds2000"])
def test_lm_teacher_forced_c4_tenth(shape):
    """configs[3] (C4: Kannala-Brandt / double-sphere, self-calibration, Cauchy, explicit Schur + dense FP64 Cholesky) at
    1/10 scale: 200 cameras (dc = 14 / 12), 100k landmarks, ~490k observations. The dense solve of the n = 2 800 system agrees
    with the oracle's to 4e-11 backward (blocked tensor-core Cholesky vs the oracle's column Cholesky on a matrix whose entries
    were summed in another order): inside the north-star 1e-9, above the 1e-11 the small cases hold."""
    teacher_forced(synth.make_shape(shape, scale=0.1), F.SCHUR_EXPLICIT, n_it=3, backward_tol=1e-10)


def test_lm_ladybug49_shape_explicit():
    """BASELINE.json configs[0]: Ladybug problem-49-7776 shape, LM + explicit Schur (direct), full size."""
    prob = synth.make_shape("ladybug49")
    g = GpuContext().upload(prob)
    rg, tg = run_lm(g, F.SCHUR_EXPLICIT, max_it=20)
    ro, to, floor = oracle_lm_with_floor(prob, F.SCHUR_EXPLICIT, max_it=20)
    assert_lm_parity(rg, tg, ro, to, floor)
    assert rg.final_cost < 0.2 * rg.initial_cost


def test_lm_trafalgar_shape_implicit_scaled():
    """BASELINE.json configs[1] shape at 1/4 scale (the oracle's single-threaded operator bounds the size)."""
    prob = synth.make_shape("trafalgar257", scale=0.25)
    g = GpuContext().upload(prob)
    rg, tg = run_lm(g, F.SCHUR_IMPLICIT, max_it=4)
    ro, to, floor = oracle_lm_with_floor(prob, F.SCHUR_IMPLICIT, max_it=4)
    assert_lm_parity(rg, tg, ro, to, floor)
    assert tg[0].ls_iter == to[0].ls_iter, "PCG iterations of the first LM iteration"


# ---- edge cases the reference's tests cover -------------------------------------------
def test_long_tracks_span_several_chunks():
    """A landmark seen by more cameras than one 256-slot tile holds (its observations span several chunks)."""
    prob = synth.make_problem(300, 3000, 6.0, seed=11, self_calibration=False)
    rng = np.random.default_rng(0)
    # give landmark 3 (at the scene centre, visible from the whole ring) an observation in every camera
    pt = prob.pt.copy(); pt[3] = 0.01
    extra_cam = np.arange(prob.ncam, dtype=np.uint32)
    keep = prob.obs_pt != 3
    obs_cam = np.concatenate([prob.obs_cam[keep], extra_cam])
    obs_pt = np.concatenate([prob.obs_pt[keep], np.full(prob.ncam, 3, np.uint32)])
    uv = np.concatenate([prob.obs_uv[keep], rng.uniform(-30, 30, (prob.ncam, 2))])
    prob2 = synth.BAProblem(camera_model=prob.camera_model, opt_flags=prob.opt_flags, pose=prob.pose, intr=prob.intr, pt=pt,
                            obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=uv, loss_id=prob.loss_id, loss_params=prob.loss_params,
                            pose_fixed=prob.pose_fixed, intr_fixed=prob.intr_fixed)
    g, o = pair(prob2)
    lam = 1e-3
    g.linearize(lam); o.linearize(lam)
    assert_blocks_close(g, o, prob2, lam, tol=1e-10)
    x = rng.standard_normal(prob2.ncam * prob2.dc)
    assert relerr(g.schur_matvec(x), o.schur_matvec(x)) < 1e-10
    for variant, tol in ((F.SCHUR_EXPLICIT, 1e-7), (F.SCHUR_IMPLICIT, 1e-5)):
        sg = g.solve_augmented(variant, lam, cg_max_iterations=1000, cg_tolerance=1e-13)
        so = o.solve_augmented(variant, lam, cg_max_iterations=1000, cg_tolerance=1e-13)
        assert relerr(sg[0], so[0]) < tol and relerr(sg[1], so[1]) < tol, variant


def test_unobserved_landmarks_and_invalid_projections():
    """Landmarks without observations (H_pp = lambda I) and points behind the camera (zero rows, Ceres convention,
    projection_factor.rs:227-239, :498-522)."""
    prob = small_problem(ncam=8, npts=200)
    # move a few landmarks behind every camera (BAL looks down -z: in front means p_cam.z < 0) by pushing them far away
    prob.pt[5] = prob.pt[5] * 1e3
    prob.pt[17] = -prob.pt[17] * 1e3
    # drop all observations of landmarks 0..3
    keep = prob.obs_pt >= 4
    prob2 = synth.BAProblem(camera_model=prob.camera_model, opt_flags=prob.opt_flags, pose=prob.pose, intr=prob.intr, pt=prob.pt,
                            obs_cam=prob.obs_cam[keep], obs_pt=prob.obs_pt[keep], obs_uv=prob.obs_uv[keep], loss_id=prob.loss_id,
                            loss_params=prob.loss_params, pose_fixed=prob.pose_fixed, intr_fixed=prob.intr_fixed)
    g, o = pair(prob2)
    assert abs(g.cost() - o.cost()) <= 1e-13 * abs(o.cost())
    g.linearize(1e-3); o.linearize(1e-3)
    rg, jcg, jpg = g.get_linearization(); ro, jco, jpo = o.get_linearization()
    assert np.array_equal(rg == 0.0, ro == 0.0), "zeroed residual rows"
    assert (ro == 0.0).all(axis=1).any(), "the case must contain invalid projections"
    assert_blocks_close(g, o, prob2, 1e-3)
    sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=300, cg_tolerance=1e-12)
    so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=300, cg_tolerance=1e-12)
    assert relerr(sg[0], so[0]) < 1e-5 and relerr(sg[1], so[1]) < 1e-5
    assert np.abs(sg[1][:4]).max() == 0.0, "unobserved landmarks do not move"


def test_fixed_variables_are_zeroed_at_update_only():
    """Problem::fix_variable: the step still contains the fixed DOF (norms include them); only the update masks them
    (src/core/problem.rs:185-289)."""
    prob = small_problem(ncam=8, npts=200, fix_first_intr=True)
    prob.pt_fixed = np.zeros(prob.npts, np.uint8)
    prob.pt_fixed[:10] = 0b101
    g, o = pair(prob)
    rg, tg = run_lm(g, F.SCHUR_EXPLICIT, max_it=3)
    ro, to = run_lm(o, F.SCHUR_EXPLICIT, max_it=3)
    _, _, floor = oracle_lm_with_floor(prob, F.SCHUR_EXPLICIT, max_it=3)
    assert_lm_parity(rg, tg, ro, to, floor)
    for a, b in zip(tg, to):
        assert abs(a.step_norm - b.step_norm) <= 1e-4 * b.step_norm
    pg, po = g.params_download(), o.params_download()
    assert relerr(pg[0][0], po[0][0]) < 1e-15, "fixed pose stays put"
    assert np.array_equal(pg[1][0], prob.intr[0]), "fixed intrinsics stay put"
    assert np.array_equal(pg[2][:10, 0], prob.pt[:10, 0]) and np.array_equal(pg[2][:10, 2], prob.pt[:10, 2])


def test_deterministic_operator_is_bitwise_reproducible(monkeypatch):
    """APEX_DETERMINISTIC=1 (what problem_upload picks by itself for a locality-ordered reconstruction): the chunk kernel flushes
    its camera windows as per-window partial rows instead of reducing into L2, and a second kernel adds each camera's rows in a
    fixed order: same bits on every run, an LM trajectory that repeats exactly, and the same operator as the reduction flush to
    summation-order level. Several camera models (dc = 6, 9, 14, 15)."""
    monkeypatch.setenv("APEX_DETERMINISTIC", "1")
    for kw in (dict(), dict(self_cal=False), dict(model=F.CAM_KANNALA_BRANDT, loss=(F.LOSS_CAUCHY, 1.0)), dict(model=F.CAM_RADTAN)):
        prob = small_problem(ncam=30, npts=2000, track=5.0, **kw)
        g, o = pair(prob)
        g.linearize(1e-3); o.linearize(1e-3)
        x = np.random.default_rng(4).standard_normal(prob.ncam * prob.dc)
        y1 = g.schur_matvec(x)
        assert np.array_equal(g.schur_matvec(x), y1) and np.array_equal(g.schur_matvec(x), y1)
        assert relerr(y1, o.schur_matvec(x)) < 1e-11
    prob = small_problem(ncam=30, npts=2000, track=5.0)
    g1, g2 = GpuContext().upload(prob), GpuContext().upload(prob)
    x = np.random.default_rng(5).standard_normal(prob.ncam * prob.dc)
    g1.linearize(1e-3)
    y_det = g1.schur_matvec(x)
    (r1, t1), (r2, t2) = run_lm(g1, F.SCHUR_IMPLICIT, max_it=4), run_lm(g2, F.SCHUR_IMPLICIT, max_it=4)
    assert [a.cost for a in t1] == [b.cost for b in t2] and r1.linear_iterations == r2.linear_iterations
    assert all(np.array_equal(a, b) for a, b in zip(g1.params_download(), g2.params_download()))
    monkeypatch.setenv("APEX_DETERMINISTIC", "0")   # windows flushed with FP64 reductions into L2 instead
    g3 = GpuContext().upload(prob)
    g3.linearize(1e-3)
    assert relerr(g3.schur_matvec(x), y_det) < 1e-12


@pytest.mark.parametrize("flush", ["0", "1"], ids=["reductions", "rows"])
def test_operator_windows_and_run_chains(flush, monkeypatch):
    """Corners of the operator kernel's camera-side bookkeeping, with both flushes: (a) three cameras - every camera's run of
    camera-sorted lanes covers several whole warps of a chunk, so the continuation chains (sums parked per warp and added after the
    next barrier) are 2-3 warps long; (b) 1 500 cameras without track locality - a range touches far more cameras than a window
    holds, so every CTA flushes and re-zeroes its window several times; (c) fewer chunks than resident CTAs."""
    monkeypatch.setenv("APEX_DETERMINISTIC", flush)
    cases = [small_problem(ncam=3, npts=3000, track=2.6, seed=5),
             small_problem(ncam=1500, npts=60000, track=5.0, window_frac=0.45, seed=6),
             small_problem(ncam=40, npts=900, track=4.0, seed=8)]
    for prob in cases:
        g, o = pair(prob)
        g.linearize(1e-3); o.linearize(1e-3)
        rng = np.random.default_rng(2)
        for _ in range(2):
            x = rng.standard_normal(prob.ncam * prob.dc)
            yg, yo = g.schur_matvec(x), o.schur_matvec(x)
            assert relerr(yg, yo) < 1e-11, (prob.ncam, relerr(yg, yo))
        sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=30, cg_tolerance=1e-6)
        so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=30, cg_tolerance=1e-6)
        assert sg[3] == so[3] and relerr(sg[0], so[0]) < 1e-5, "truncated PCG: same iteration count, same step"


@pytest.mark.parametrize("tail", ["2", "1", "0", "stride", "stride-nopdl", "nopdl"])
def test_pcg_fused_tail_and_three_kernel_path(tail, monkeypatch):
    """PCG between two operator applications: the fused tail kernel (one warp per camera, tagged-slot sum exchanges, second pass of
    the deterministic flush inside; APEX_PCG_TAIL = CTAs per SM), its grid-stride variant (the one problems with more cameras than
    16 x SMs get - the Final-13682 shape -, forced here by APEX_PCG_TAIL_STRIDE with about three cameras per warp), both with and
    without programmatic dependent launches (APEX_PDL=0), and the separate kernels (APEX_PCG_TAIL=0) take the oracle's iteration
    count and agree with it on the step; ncam = 37 leaves warps of the last CTA without cameras."""
    if tail.startswith("stride"):
        monkeypatch.setenv("APEX_PCG_TAIL_STRIDE", "1")
    if tail.endswith("nopdl"):
        monkeypatch.setenv("APEX_PDL", "0")
    if tail in ("2", "1", "0"):
        monkeypatch.setenv("APEX_PCG_TAIL", tail)
    for ncam, npts in ((37, 1500), (10, 300)):
        prob = small_problem(ncam=ncam, npts=npts)
        g, o = pair(prob)
        sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, cg_max_iterations=40, cg_tolerance=1e-6)
        so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, cg_max_iterations=40, cg_tolerance=1e-6)
        assert sg[3] == so[3], "PCG iterations"
        sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, cg_max_iterations=600, cg_tolerance=1e-12)
        so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-2, cg_max_iterations=600, cg_tolerance=1e-12)
        assert relerr(sg[0], so[0]) < 1e-6 and relerr(sg[1], so[1]) < 1e-6
    (rg, tg) = run_lm(GpuContext().upload(prob), F.SCHUR_IMPLICIT, max_it=5)
    (ro, to) = run_lm(OracleContext().upload(prob), F.SCHUR_IMPLICIT, max_it=5)
    assert (rg.status, rg.iterations, [t.accepted for t in tg]) == (ro.status, ro.iterations, [t.accepted for t in to])


def test_observer_feed():
    """f4: the observer feed of the LM loop (levenberg_marquardt.rs:930-940, :1010-1011) through the C ABI's callbacks, same contract as
    the oracle's (parity_helpers.check_observer_feed); the first iteration's tuple against the oracle's at the LM tolerances."""
    from parity_helpers import DIRECT_FORWARD_BLUNDER, PCG_FORWARD_BLUNDER, check_observer_feed, rel
    prob = small_problem(ncam=10, npts=300, seed=11)
    for variant in (F.SCHUR_EXPLICIT, F.SCHUR_IMPLICIT):
        sg, rg = check_observer_feed(GpuContext, prob, variant)
        so, ro = check_observer_feed(OracleContext, prob, variant)
        assert (rg.status, rg.iterations) == (ro.status, ro.iterations)
        assert [s[1] for s in sg] == [s[1] for s in so], "accept pattern"
        # forward agreement after one full iteration: the blunder bounds of the teacher-forced harness (parity_helpers: the solve is only
        # determined to cond(S) * eps, truncated PCG further out); the gradient norm at the start is a well-conditioned quantity
        g0, o0 = sg[0], so[0]
        fwd = DIRECT_FORWARD_BLUNDER if variant == F.SCHUR_EXPLICIT else PCG_FORWARD_BLUNDER
        assert rel(g0[2], o0[2]) < fwd and rel(g0[3], o0[3]) < 1e-11 and rel(g0[5], o0[5]) < 100 * fwd


def test_bal_file_to_gpu_solve(tmp_path):
    """Data format either side of the path: generator -> BAL text -> apex_bal_load -> apex_bal_build_problem (the CLI's
    construction, bin/bundle_adjustment.rs:212-441) -> GPU LM, against the oracle on the same loaded problem; and the
    CLI binary on the same file reports the same iteration count and costs."""
    import os
    import re
    import subprocess
    from apex_solver_b200.bal import dataset_from_problem, load_bal
    path = str(tmp_path / "problem-12-400-pre.txt")
    dataset_from_problem(small_problem(ncam=12, npts=400, seed=41)).write(path)
    prob = load_bal(path).problem(optimization_type="bundle-adjustment")
    g, o = pair(prob)
    (rg, tg), (ro, to) = run_lm(g, F.SCHUR_IMPLICIT, max_it=20), run_lm(o, F.SCHUR_IMPLICIT, max_it=20)
    assert (rg.status, rg.iterations) == (ro.status, ro.iterations)
    assert abs(rg.final_cost - ro.final_cost) <= FINAL_RTOL * abs(ro.final_cost)
    exe = os.path.join(os.path.dirname(F.LIB_PATH), "bundle_adjustment")
    # `-s matrix-free` = APEX_SCHUR_IMPLICIT; the default `-s implicit` is what the reference binary dispatches to today:
    # explicit S + scalar-Jacobi PCG (explicit_schur.rs:1222-1225) = APEX_SCHUR_EXPLICIT_PCG
    # Both paths are bitwise reproducible (deterministic flush of the operator's windows; S formed block by block in a fixed order):
    # the CLI prints the library's number (7 digits).
    for flags, variant, tol in ((["-s", "matrix-free"], F.SCHUR_IMPLICIT, 1e-6), ([], F.SCHUR_EXPLICIT_PCG, 1e-6)):
        rv, _ = run_lm(GpuContext().upload(prob), variant, max_it=20)
        r = subprocess.run([exe, path, "-t", "bundle-adjustment", "-v"] + flags, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert int(re.search(r"Iterations: (\d+)", r.stdout).group(1)) == rv.iterations
        assert abs(float(re.search(r"Final cost: (\S+)", r.stdout).group(1)) - rv.final_cost) <= tol * abs(rv.final_cost)


def test_error_behaviour():
    g = GpuContext()
    with pytest.raises(F.ApexError) as e:
        g.cost()
    assert e.value.status == F.ERR_INVALID_STATE
    prob = small_problem(ncam=6, npts=60)
    g.upload(prob)
    with pytest.raises(F.ApexError) as e:
        g.schur_matvec(np.zeros(prob.ncam * prob.dc))
    assert e.value.status == F.ERR_INVALID_STATE  # not linearized
    bad = small_problem(ncam=6, npts=60)
    bad.obs_cam = bad.obs_cam.copy(); bad.obs_cam[0] = 99
    with pytest.raises(F.ApexError) as e:
        GpuContext().upload(bad)
    assert e.value.status == F.ERR_INVALID_INPUT
    cfg = g.default_config(True)
    cfg.compute_covariances = 1   # accepted without effect: the reference's Schur solvers return None (src/linalg/mod.rs:170-172)
    cfg.max_iterations = 1
    res, _ = g.lm_solve(cfg)
    assert res.iterations == 2


@pytest.mark.parametrize("variant,kw", [(F.SCHUR_EXPLICIT, {}), (F.SCHUR_IMPLICIT, dict(cg_it=3000, cg_tolerance=1e-12))], ids=["explicit", "implicit_converged"])
@pytest.mark.parametrize("self_cal", [True, False], ids=["selfcal", "ba"])
def test_lm_with_jacobi_scaling(variant, kw, self_cal):
    """LevenbergMarquardtConfig::with_jacobi_scaling(true) (levenberg_marquardt.rs:871-879, optimizer/mod.rs:749-763): column
    scaling 1 / (1 + ||J column||) fixed at the first iterate, solve in scaled variables, step scaled back, predicted reduction from
    the unscaled step and the scaled gradient. Same accept pattern / status / iteration count as the oracle, per-iteration cost
    and final parameters at the usual tolerances; and the scaling must matter (another trajectory than without it)."""
    prob = small_problem(ncam=14, npts=500, self_cal=self_cal)
    g, o = pair(prob)
    (rg, tg), (ro, to) = run_lm(g, variant, max_it=6, use_jacobi_scaling=1, **kw), run_lm(o, variant, max_it=6, use_jacobi_scaling=1, **kw)
    assert (rg.status, rg.iterations, [t.accepted for t in tg]) == (ro.status, ro.iterations, [t.accepted for t in to])
    for a, b in zip(tg, to):
        assert abs(a.cost - b.cost) <= 1e-8 * abs(b.cost) and abs(a.gradient_norm - b.gradient_norm) <= 1e-8 * b.gradient_norm
        assert abs(a.step_norm - b.step_norm) <= 1e-6 * b.step_norm
    for xa, xb in zip(g.params_download(), o.params_download()):
        assert relerr(xa, xb) < 1e-6
    (r0, t0) = run_lm(GpuContext().upload(prob), variant, max_it=6, **kw)
    assert abs(t0[0].gradient_norm - tg[0].gradient_norm) > 1e-3 * t0[0].gradient_norm, "scaled gradient norm is reported"
    g.params_upload(*o.params_download())
    g.linearize(1e-3); o.linearize(1e-3)   # the standalone entry points stay unscaled after a scaled solve
    assert relerr(g.get_linearization()[1], o.get_linearization()[1]) < 1e-12


def test_empty_and_degenerate_problems():
    """Edge cases of the reference's front door: a problem without residual blocks is OptimizerError::NoResidualBlocks
    (src/optimizer/mod.rs:66-141), a single observation / a camera without observations / ragged last chunks solve."""
    base = small_problem(ncam=5, npts=40)
    empty = BAProblemLike(base, np.zeros(0, bool))
    for ctx in (GpuContext().upload(empty), OracleContext().upload(empty)):
        with pytest.raises(F.ApexError) as e:
            run_lm(ctx, F.SCHUR_IMPLICIT)
        assert e.value.status == F.ERR_NO_RESIDUAL_BLOCKS
    keep = np.zeros(base.nobs, bool); keep[0] = True
    one = BAProblemLike(base, keep)
    (rg, tg), (ro, to) = run_lm(GpuContext().upload(one), F.SCHUR_EXPLICIT, max_it=3), run_lm(OracleContext().upload(one), F.SCHUR_EXPLICIT, max_it=3)
    assert (rg.status, rg.iterations) == (ro.status, ro.iterations) and rg.final_cost < rg.initial_cost
    assert abs(rg.initial_cost - ro.initial_cost) <= 1e-13 * ro.initial_cost
    assert abs(rg.final_cost - ro.final_cost) <= 1e-3 * ro.final_cost   # 1 observation, 3 + 15 + 120 unknowns: everything but 2 directions is held by lambda alone
    nocam = BAProblemLike(base, base.obs_cam != 3)          # camera 3 keeps its variables but sees nothing
    g, o = pair(nocam)
    g.linearize(1e-3); o.linearize(1e-3)
    x = np.random.default_rng(2).standard_normal(nocam.ncam * nocam.dc)
    assert relerr(g.schur_matvec(x), o.schur_matvec(x)) < 1e-10
    (rg, tg), (ro, to) = run_lm(g, F.SCHUR_IMPLICIT, max_it=4), run_lm(o, F.SCHUR_IMPLICIT, max_it=4)
    assert (rg.status, rg.iterations) == (ro.status, ro.iterations)


def test_high_degree_cameras_and_chunk_boundaries():
    """Cameras with more observations than one camera work item holds (2 048): H_cc / g_c / Schur-Jacobi partial sums are
    combined across work items; landmarks with exactly 1, 255, 256 and 257 observations sit on the chunk boundaries of the
    point-major layout (257 is a multi-chunk landmark)."""
    base = small_problem(ncam=4, npts=7000, track=3.0, seed=5)      # every camera sees ~5 000 landmarks
    assert np.bincount(base.obs_cam).min() > 2048
    g, o = pair(base)
    g.linearize(1e-3); o.linearize(1e-3)
    assert_blocks_close(g, o, base, 1e-3)
    sg = g.solve_augmented(F.SCHUR_EXPLICIT, 1e-3)
    so = o.solve_augmented(F.SCHUR_EXPLICIT, 1e-3)
    assert relerr(sg[0], so[0]) < 1e-7 and relerr(sg[1], so[1]) < 1e-7
    sg = g.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=500, cg_tolerance=1e-10)   # 36 camera dofs: PCG fights rounding, counts are not comparable
    so = o.solve_augmented(F.SCHUR_IMPLICIT, 1e-3, cg_max_iterations=500, cg_tolerance=1e-10)
    assert relerr(sg[0], so[0]) < 1e-4 and relerr(sg[1], so[1]) < 1e-4
    # track lengths on the chunk boundaries: 300 cameras, landmarks 0..3 seen by exactly 1 / 255 / 256 / 257 of them
    prob = small_problem(ncam=300, npts=400, track=4.0, seed=6)
    keep = prob.obs_pt >= 4
    cams, pts = [], []
    for lp, k in enumerate((1, 255, 256, 257)):
        cams.append(np.arange(k, dtype=np.uint32)); pts.append(np.full(k, lp, np.uint32))
    extra_cam, extra_pt = np.concatenate(cams), np.concatenate(pts)
    from apex_solver_b200.context import BAProblem
    p2 = BAProblem(camera_model=prob.camera_model, opt_flags=prob.opt_flags, pose=prob.pose, intr=prob.intr, pt=prob.pt,
                   obs_cam=np.concatenate([prob.obs_cam[keep], extra_cam]), obs_pt=np.concatenate([prob.obs_pt[keep], extra_pt]),
                   obs_uv=np.concatenate([prob.obs_uv[keep], np.random.default_rng(1).uniform(-50, 50, (extra_cam.size, 2))]),
                   loss_id=prob.loss_id, loss_params=prob.loss_params, pose_fixed=prob.pose_fixed)
    g, o = pair(p2)
    assert abs(g.cost() - o.cost()) <= 1e-13 * abs(o.cost())
    g.linearize(1e-3); o.linearize(1e-3)
    assert_blocks_close(g, o, p2, 1e-3)
    x = np.random.default_rng(3).standard_normal(p2.ncam * p2.dc)
    assert relerr(g.schur_matvec(x), o.schur_matvec(x)) < 1e-10


def BAProblemLike(p, keep):
    from apex_solver_b200.context import BAProblem
    return BAProblem(camera_model=p.camera_model, opt_flags=p.opt_flags, pose=p.pose, intr=p.intr, pt=p.pt, obs_cam=p.obs_cam[keep], obs_pt=p.obs_pt[keep],
                     obs_uv=p.obs_uv[keep], loss_id=p.loss_id, loss_params=p.loss_params, intr_vars_present=p.intr_vars_present, pose_fixed=p.pose_fixed)


def test_kernels_ran_on_the_device():
    prob = small_problem(ncam=6, npts=60)
    g = GpuContext().upload(prob)
    n0 = g.kernel_launches()
    g.linearize(1e-3)
    g.solve_augmented(F.SCHUR_IMPLICIT, 1e-3)
    assert g.kernel_launches() > n0 + 5


# ---- size-independent properties at the full BASELINE.json shape -----------------------------------------
def test_schur_operator_properties_trafalgar_full():
    """configs[1] at full size (257 cams / 65k pts / ~226k obs): S is symmetric positive definite and linear;
    S x = b solved by PCG reproduces b; the oracle agrees on the operator."""
    prob = synth.make_shape("trafalgar257")
    g = GpuContext().upload(prob)
    lam = 1e-3
    g.linearize(lam)
    n = prob.ncam * prob.dc
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Sx, Sy = g.schur_matvec(x), g.schur_matvec(y)
    assert abs(y @ Sx - x @ Sy) <= 1e-10 * abs(y @ Sx), "symmetry"
    assert x @ Sx > 0 and y @ Sy > 0, "positive definite"
    assert relerr(g.schur_matvec(2.0 * x - 3.0 * y), 2.0 * Sx - 3.0 * Sy) < 1e-11, "linearity"
    o = OracleContext().upload(prob)
    o.linearize(lam)
    assert relerr(Sx, o.schur_matvec(x)) < 1e-10
    # explicit and implicit variants solve the same system
    s_imp = g.solve_augmented(F.SCHUR_IMPLICIT, lam, cg_max_iterations=2000, cg_tolerance=1e-13)
    s_exp = g.solve_augmented(F.SCHUR_EXPLICIT, lam)
    assert relerr(s_imp[0], s_exp[0]) < 1e-5 and relerr(s_imp[1], s_exp[1]) < 1e-5


def test_operator_and_lm_properties_venice_full():
    """The bench workload (configs[2], Venice-1778 shape at full size: 1 778 cams / 994k pts / 5.3 M obs) is too large for
    the oracle; checked through size-independent properties: S symmetric, positive definite, linear; the default and the
    deterministic operator agree; the reduced system solved by PCG satisfies S dx = b to the CG tolerance
    (recomputed with the operator); an accepted LM step lowers the cost, and the cost the LM loop reports equals the cost
    kernel's value at the downloaded parameters."""
    import os
    prob = synth.make_shape("venice1778")
    g = GpuContext().upload(prob)
    lam = 1e-3
    g.linearize(lam)
    n = prob.ncam * prob.dc
    rng = np.random.default_rng(8)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Sx, Sy = g.schur_matvec(x), g.schur_matvec(y)
    assert abs(y @ Sx - x @ Sy) <= 1e-10 * abs(y @ Sx), "symmetry"
    assert x @ Sx > 0 and y @ Sy > 0, "positive definite"
    assert relerr(g.schur_matvec(2.0 * x - 3.0 * y), 2.0 * Sx - 3.0 * Sy) < 1e-11, "linearity"
    os.environ["APEX_DETERMINISTIC"] = "1"   # the deterministic two-pass operator agrees with the default one
    try:
        g2 = GpuContext().upload(prob)
        g2.linearize(lam)
        y2 = g2.schur_matvec(x)
        assert relerr(y2, Sx) < 1e-12 and np.array_equal(g2.schur_matvec(x), y2)
        g2.close()
    finally:
        del os.environ["APEX_DETERMINISTIC"]
    cfg = g.default_config(True)
    cfg.schur_variant = F.SCHUR_IMPLICIT
    cfg.max_iterations = 6
    res, tr = g.lm_solve(cfg)
    assert res.iterations >= 1 and all(np.isfinite(t.cost) for t in tr)
    costs = [res.initial_cost] + [t.cost for t in tr]
    for prev, t in zip(costs, tr):
        assert (t.cost < prev) if t.accepted else (t.cost == prev), "accepted steps lower the cost, rejected ones keep it"
    assert any(t.accepted for t in tr) and res.final_cost < res.initial_cost
    assert abs(g.cost() - res.final_cost) <= 1e-12 * res.final_cost, "reported cost = cost kernel at the final parameters"

"""Teacher-forced single-iteration parity.

A free-running LM comparison lets rounding differences random-walk along the trajectory: after a few iterations on a
self-calibrating problem (cond(S) ~1e10..1e12) two runs of the SAME algorithm with another summation order differ by
1e-6 and more. Teacher forcing removes that: at every iterate of the ORACLE's trajectory the device under test (DUT) is
given the oracle's parameters, damping and nu and runs exactly ONE Levenberg-Marquardt iteration
(levenberg_marquardt.rs:770-817, optimizer/mod.rs:591-675) with the linear solve converged.

Even one iteration cannot be compared FORWARD at 1e-9: measured on the CPU (oracle against itself in other summation
orders), the trial cost of a single iteration moves by up to 3e-6 because the linear solve is only determined to
cond(S) * eps and PCG leaves through the reference's absolute breakdown test |p.Ap| < 1e-20 (implicit_schur.rs:626-629)
20 iterations earlier or later depending on rounding. So the iteration is taken apart into links that ARE well
conditioned, each checked against the oracle's arithmetic on the DUT's own inputs (a = DUT, b = oracle):

  1. cost at the iterate                    a vs b                                           <= 1e-13 relative
  2. gradient norm ||J^T r||                a vs b                                           <= 1e-12
  3. linear solve, camera step              ||S (dc_a - dc_b)|| <= BACKWARD_TOL ||S dc_b||   (S applied by the oracle:
                                            "both solved the same reduced system"; forward differences are amplified by
                                            cond(S), their image under S is not) - 1e-11 direct, 1e-9 PCG, or 10 x what
                                            the oracle and its twin reach where PCG leaves through the breakdown test
  4. back-substitution                      dp_a vs oracle back-substitution OF dc_a         <= 1e-9
  5. step norm, predicted reduction         a vs numpy on (dc_a, dp_a, oracle gradient)      <= 1e-12 / 1e-9
  6. trial cost                             a vs oracle cost at  x (+) step_a                <= 1e-12
  7. rho, damping, nu, accept flag          a vs the oracle's compute_step_quality / update_damping on a's own numbers
                                                                                             <= 1e-14 / exact
  8. parameters after accept or revert      a vs oracle apply_step(+-step_a)                 <= 1e-13
  9. forward, a vs b (trial cost, step norm, rho): REPORTED in the returned rows next to what the oracle's own twin shows
     at the same iterate (measured on B200: 1e-12..1e-8 with the direct solver, where one reversed-order twin
     underestimates the floor by 10-20x: the GPU differs from the oracle in the Cholesky blocking and the triangular
     solves as well, not only in the order S is summed in); asserted only as a blunder bound on accepted steps
     (1e-6 direct, 1e-4 PCG).
Links 1-8 hold the DUT to the reference arithmetic at <= 1e-9 everywhere the arithmetic is well conditioned; link 3 is
the conditioning-free statement about the solve.
"""
import ctypes as C

import numpy as np

from apex_solver_b200 import _ffi as F

NORTH_STAR = 1e-9
FLOOR_FACTOR = 10.0
TOL_CAP = 1e-5
PCG_FORWARD_BLUNDER = 1e-4
DIRECT_FORWARD_BLUNDER = 1e-6
BACKWARD_TOL = {F.SCHUR_EXPLICIT: 1e-11, F.SCHUR_IMPLICIT: 1e-9, F.SCHUR_EXPLICIT_PCG: 1e-9}


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    return float(np.abs(a - b).max(initial=0.0) / scale)


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def one_iteration(ctx, variant, lam, nu, cg_it, cg_tol, precond=F.PRECOND_SCHUR_JACOBI):
    cfg = ctx.default_config(True)
    cfg.schur_variant = variant
    cfg.schur_preconditioner = precond
    cfg.max_iterations = 0          # iterations = max_iterations + 1 (levenberg_marquardt.rs:1015)
    cfg.cg_max_iterations = cg_it
    cfg.cg_tolerance = cg_tol
    cfg.damping, cfg.damping_nu = lam, nu
    res, tr = ctx.lm_solve(cfg)
    assert res.iterations == 1
    return res, tr[0], cfg


def normalized_poses(lib, pose):
    """SE3::from(DVector) (se3.rs:200-206), what lm_solve does to the initial poses, with the oracle's arithmetic."""
    out = np.empty_like(pose)
    for i in range(pose.shape[0]):
        lib.oracle_se3_normalize(F.ptr(pose[i]), out[i].ctypes.data)
    return out


def teacher_forced_parity(prob, variant, dut, oracle, scratch, lib, twin=None, n_it=6, cg_it=3000, cg_tol=1e-12, lam0=1e-3, nu0=2.0, log=None, backward_tol=None):
    """dut / oracle / scratch (/ twin): contexts with `prob` uploaded; oracle walks its own LM trajectory, scratch (another
    oracle context) recomputes links 3-8 on the DUT's numbers, twin (the oracle in another summation order) gives the
    forward floor of link 9 for the direct solver and the backward floor of link 3 for PCG. lib = the oracle library (scalar LM functions).
    Returns one report row per iterate."""
    lam, nu = lam0, nu0
    report = []
    unref_intr = bool(prob.intr_vars_present) and not (prob.opt_flags & F.OPT_INTRINSIC)
    for it in range(n_it):
        pose, intr, pt = oracle.params_download()
        x = (normalized_poses(lib, pose), intr, pt)
        for ctx in (dut, oracle, scratch) + ((twin,) if twin is not None else ()):
            ctx.params_upload(*x)
        ra, a, cfg = one_iteration(dut, variant, lam, nu, cg_it, cg_tol)
        dca, dpa = dut.get_step()
        pa = dut.params_download()
        rb, b, _ = one_iteration(oracle, variant, lam, nu, cg_it, cg_tol)
        dcb, dpb = oracle.get_step()
        tag = f"iterate {it} (lambda {lam:.3e})"
        # 1, 2
        assert rel(ra.initial_cost, rb.initial_cost) <= 1e-13, f"{tag}: cost {ra.initial_cost!r} vs {rb.initial_cost!r}"
        assert rel(a.gradient_norm, b.gradient_norm) <= 1e-12, f"{tag}: gradient norm {a.gradient_norm!r} vs {b.gradient_norm!r}"
        # 3: backward agreement of the camera steps
        scratch.linearize(lam)
        s_diff = scratch.schur_matvec((dca - dcb).ravel())
        s_ref = scratch.schur_matvec(dcb.ravel())
        backward = float(np.linalg.norm(s_diff) / max(np.linalg.norm(s_ref), 1e-300))
        tol_backward = BACKWARD_TOL[variant] if backward_tol is None else backward_tol
        f = None
        if twin is not None:
            _, f, _ = one_iteration(twin, variant, lam, nu, cg_it, cg_tol)
            # PCG may leave through the reference's ABSOLUTE breakdown test |p.Ap| < 1e-20 before the residual is at
            # 1e-9 |b| (badly scaled S): the oracle against its own twin shows how far this iterate can be solved. The
            # direct solver's backward error grows with the size of S (C4 at 1/10 scale, n = 2 800: 4e-11); the twin bounds it
            # the same way, but never beyond the north-star 1e-9.
            dct, _ = twin.get_step()
            floor_b = float(np.linalg.norm(scratch.schur_matvec((dct - dcb).ravel())) / max(np.linalg.norm(s_ref), 1e-300))
            tol_backward = max(tol_backward, FLOOR_FACTOR * floor_b)
            cap = NORTH_STAR if variant == F.SCHUR_EXPLICIT else 1e-6
            assert tol_backward <= cap, f"{tag}: the oracle's own solve only reproduces to {floor_b:.1e} backward"
        assert backward <= tol_backward, f"{tag}: ||S (dc_dut - dc_oracle)|| / ||S dc_oracle|| = {backward:.2e} (tolerance {tol_backward:.1e})"
        # 4: back-substitution of the DUT's own camera step
        dp_ref = scratch.back_substitute(dca, implicit_flavour=(variant == F.SCHUR_IMPLICIT))
        e_back = relerr(dpa, dp_ref)
        assert e_back <= 1e-9, f"{tag}: back-substitution {e_back:.2e}"
        # 5: norms on the DUT's own step with the oracle's gradient (+J^T r)
        _, gc, _, gp, _ = scratch.get_blocks()
        s2 = float((dca * dca).sum() + (dpa * dpa).sum())
        sg = float((dca * gc).sum() + (dpa * gp).sum())
        assert rel(a.step_norm, np.sqrt(s2)) <= 1e-12, f"{tag}: step norm"
        pred = 0.5 * (lam * a.step_norm * a.step_norm - sg)   # compute_predicted_reduction (levenberg_marquardt.rs:721-727)
        assert rel(a.predicted_reduction, pred) <= 1e-9, f"{tag}: predicted reduction {a.predicted_reduction!r} vs {pred!r}"
        # 6: trial cost at x (+) step_a
        scratch.apply_step(dca, dpa, +1.0)
        trial = scratch.cost()
        # (a wild rejected step lands where the cost is 1e8 x larger and every ulp of the parameters shows in it)
        assert rel(a.new_cost, trial) <= (1e-12 if a.new_cost <= 10.0 * ra.initial_cost else 1e-9), f"{tag}: trial cost {a.new_cost!r} vs oracle at the DUT's trial point {trial!r}"
        # 7: scalar bookkeeping on the DUT's own numbers
        rho = lib.oracle_compute_step_quality(ra.initial_cost, a.new_cost, a.predicted_reduction)
        assert abs(a.tr_ratio - rho) <= 1e-14 * max(abs(rho), 1.0), f"{tag}: rho {a.tr_ratio!r} vs {rho!r}"
        d, v = C.c_double(lam), C.c_double(nu)
        accepted = lib.oracle_update_damping(C.byref(d), C.byref(v), cfg.damping_min, cfg.damping_max, a.tr_ratio)
        assert a.accepted == accepted, f"{tag}: accept flag"
        assert rel(a.tr_radius, d.value) <= 1e-14 and ra.final_damping_nu == v.value, f"{tag}: damping {a.tr_radius!r} vs {d.value!r}, nu {ra.final_damping_nu} vs {v.value}"
        assert rel(a.cost, a.new_cost if accepted else ra.initial_cost) == 0.0, f"{tag}: current cost after the iteration"
        # 8: parameters after accept / revert
        if not accepted:
            scratch.apply_step(dca, dpa, -1.0)
        for xa, xs, what in zip(pa, scratch.params_download(), ("poses", "intrinsics", "landmarks")):
            if what == "intrinsics" and unref_intr:
                continue  # unreferenced intr_* variables: zero step (checked by the parameter norm below)
            assert relerr(xa, xs) <= 1e-13, f"{tag}: {what} after {'accept' if accepted else 'revert'}: {relerr(xa, xs):.2e}"
        # 9: forward comparison with the oracle's own iteration (reported; blunder bound on accepted steps)
        fwd = {f: rel(getattr(a, f), getattr(b, f)) for f in ("new_cost", "step_norm", "predicted_reduction")}
        fwd["rho_abs"] = abs(a.tr_ratio - b.tr_ratio)
        tol_fwd = rel(f.new_cost, b.new_cost) if f is not None else None   # what the oracle's own twin shows at this iterate
        if b.accepted:  # a rejected trial point can sit anywhere (cost 1e8 x the current one): nothing to compare forward
            bound = DIRECT_FORWARD_BLUNDER if variant == F.SCHUR_EXPLICIT else PCG_FORWARD_BLUNDER
            assert fwd["new_cost"] <= bound, f"{tag}: trial cost {a.new_cost!r} vs {b.new_cost!r}"
        row = dict(iterate=it, accepted=int(b.accepted), lam=lam, backward=backward, back_sub=e_back, ls_iter=(int(a.ls_iter), int(b.ls_iter)),
                   fwd_new_cost=fwd["new_cost"], fwd_step_norm=fwd["step_norm"], fwd_rho=fwd["rho_abs"], tol_fwd=tol_fwd,
                   same_accept=bool(a.accepted == b.accepted))
        report.append(row)
        if log:
            log(row)
        lam, nu = rb.final_damping, rb.final_damping_nu
    return report


def check_observer_feed(ctx_factory, prob, variant=F.SCHUR_EXPLICIT, max_it=6):
    """The observer contract of the LM loop (OptObserver, src/observers/mod.rs:201-330; fed at levenberg_marquardt.rs:930-940 and
    :1010-1011), shared by the oracle test here and the GPU test: one on_step per iteration carrying the tuple of
    notify_observers_generic - which is what the IterationStats row of that iteration holds (cost, gradient norm, tr_radius = damping,
    step norm, tr_ratio = rho) -, iterations numbered from 0, the variables readable from inside the callback, exactly one
    on_optimization_complete with SolverResult::iterations (levenberg_marquardt.rs:1642-1682), observers kept across solves until
    cleared, several observers fed in registration order."""
    ctx = ctx_factory().upload(prob)
    steps, done, order, params_seen = [], [], [], []

    def on_step(c, m):
        steps.append((m.iteration, m.accepted, m.cost, m.gradient_norm, m.damping, m.step_norm, m.step_quality))
        params_seen.append([a.copy() for a in c.params_download()])
        order.append("a")

    ctx.add_observer(on_step, lambda c, n: done.append(n))
    ctx.add_observer(lambda c, m: order.append("b"))          # a second observer without a completion hook
    cfg = ctx.default_config(True)
    cfg.schur_variant = variant
    cfg.max_iterations = max_it
    res, trace = ctx.lm_solve(cfg)
    assert done == [res.iterations], "on_optimization_complete exactly once, with the result's iteration count"
    assert [s[0] for s in steps] == list(range(res.iterations))
    assert order == ["a", "b"] * res.iterations
    for s, t in zip(steps, trace):
        assert (s[1], s[2], s[3], s[4], s[5], s[6]) == (t.accepted, t.cost, t.gradient_norm, t.tr_radius, t.step_norm, t.tr_ratio)
    final = ctx.params_download()
    for a, b in zip(params_seen[-1], final):
        assert np.array_equal(a, b), "the values an observer reads at the last step are the result's"
    if steps[0][1]:   # the first step was accepted: the observer already sees the moved variables (on_step comes after the update)
        assert any(not np.array_equal(a, b) for a, b in zip(params_seen[0], (prob.pose, prob.intr, prob.pt)))
    ctx.upload(prob)                                         # observers survive a new upload and a second solve ...
    ctx.lm_solve(cfg)
    assert len(done) == 2 and len(steps) == 2 * res.iterations
    ctx.clear_observers()                                    # ... until they are cleared
    ctx.upload(prob)
    ctx.lm_solve(cfg)
    assert len(done) == 2 and len(steps) == 2 * res.iterations
    return steps[: res.iterations], res

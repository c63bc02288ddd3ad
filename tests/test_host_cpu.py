"""CPU-only tests (no GPU in this container): the C-ABI library loads and exports every symbol include/apex_gpu.h
declares, fails loudly without a device (no CPU fallback), the host-side sharding logic, the synthetic generator,
the oracle end to end on small BASELINE-shaped problems, and the N>1 decomposition over gloo (world_size 2)."""
import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np
import dataclasses

import pytest

from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import BAProblem, GpuContext, layout_stats, shard_info
from oracle_backend import OracleContext, oracle_lib
from parity_helpers import check_observer_feed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_GPU = F.load_library().apex_device_count() > 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "apex_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(apex_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    lib = C.CDLL(F.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/apex_gpu.h but not exported"
    assert {"apex_" + n for n in F.SYMBOLS} == declared, "the ctypes table and the header disagree"
    assert lib.apex_abi_version() == 100


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors vs the C compiler's view of include/apex_gpu.h (sizes and a few offsets)."""
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "apex_gpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(apex_ctx_desc),sizeof(apex_problem_desc),sizeof(apex_lm_config),sizeof(apex_lm_result),sizeof(apex_iter_trace),"
                   "sizeof(apex_dims),sizeof(apex_profile),sizeof(apex_layout_stats),offsetof(apex_problem_desc,loss_params),offsetof(apex_lm_config,cg_tolerance),"
                   "offsetof(apex_lm_result,linear_iterations),sizeof(apex_observer_metrics),sizeof(apex_observer),offsetof(apex_observer_metrics,step_quality));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(F.CtxDesc), C.sizeof(F.ProblemDesc), C.sizeof(F.LmConfig), C.sizeof(F.LmResult), C.sizeof(F.IterTrace), C.sizeof(F.Dims),
            C.sizeof(F.Profile), C.sizeof(F.LayoutStats), F.ProblemDesc.loss_params.offset, F.LmConfig.cg_tolerance.offset, F.LmResult.linear_iterations.offset,
            C.sizeof(F.ObserverMetrics), C.sizeof(F.Observer), F.ObserverMetrics.step_quality.offset]
    assert got == want


def test_plain_c_client_links_and_calls_the_host_side_of_the_abi(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (no C++, no Python) includes include/apex_gpu.h, links libapex_gpu.so and calls
    the entry points that need no device - presets, the sharding rule, the BAL round trip (writer -> loader -> the CLI's problem
    construction) and the layout build - the way the reference's Rust shim would through `extern "C"`. Without a device
    apex_ctx_create must report APEX_ERR_NO_DEVICE (no CPU fallback); with one it must succeed."""
    src = tmp_path / "client.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "apex_gpu.h"
int main(int argc, char** argv) {
  if (apex_abi_version() != 100) return 10;
  apex_lm_config cfg;
  apex_lm_config_for_bundle_adjustment(&cfg);
  if (cfg.max_iterations != 20 || cfg.damping != 1e-3) return 11;
  /* 2 cameras, 3 points, 5 observations */
  double cams[2 * 9] = {0.01, -0.02, 0.03, 0.1, 0.2, -0.3, 500.0, -1e-7, 1e-13,   -0.02, 0.01, 0.0, -0.1, 0.0, 0.2, 510.0, 2e-7, 0.0};
  double pts[3 * 3] = {0.1, 0.2, -5.0, -0.3, 0.1, -6.0, 0.2, -0.2, -4.0};
  unsigned obs_cam[5] = {0, 0, 0, 1, 1}, obs_pt[5] = {0, 1, 2, 0, 2};
  double uv[5 * 2] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
  apex_bal_dataset* ds = 0;
  if (apex_bal_from_arrays(2, 3, 5, cams, pts, obs_cam, obs_pt, uv, &ds) != APEX_OK) return 12;
  if (apex_bal_write(ds, argv[1]) != APEX_OK) return 13;
  apex_bal_free(ds);
  if (apex_bal_load(argv[1], &ds) != APEX_OK) { fprintf(stderr, "%s\n", apex_bal_last_error()); return 14; }
  apex_bal_view v;
  if (apex_bal_view_get(ds, &v) != APEX_OK || v.ncam != 2 || v.npts != 3 || v.nobs != 5) return 15;
  if (memcmp(v.cameras, cams, sizeof cams) != 0 || memcmp(v.points, pts, sizeof pts) != 0) return 16;   /* 17 significant digits: bit-exact */
  apex_problem_desc desc;
  if (apex_bal_build_problem(ds, 3 /* -n: all points */, 1 /* SelfCalibration */, &desc) != APEX_OK) return 17;
  if (desc.ncam != 2 || desc.npts != 3 || desc.nobs != 5 || desc.intr_dim != 3) return 18;
  unsigned block = 0, npl = 0; unsigned long long nl = 0;
  if (apex_shard_info(desc.npts, desc.nobs, desc.obs_pt, 2, 1, &block, &npl, (uint64_t*)&nl) != APEX_OK || block != 128 || npl != 0 || nl != 0) return 19;
  apex_layout_stats st;
  if (apex_layout_stats_compute(&desc, 1, 0, &st) != APEX_OK || st.consistent != 1 || st.nobs_local != 5 || st.nchunks != 1) return 20;
  apex_ctx_desc cd; memset(&cd, 0, sizeof cd); cd.nranks = 1;
  apex_ctx* ctx = 0;
  apex_status s = apex_ctx_create(&cd, &ctx);
  if (apex_device_count() > 0) { if (s != APEX_OK) return 21; apex_ctx_destroy(ctx); }
  else if (s != APEX_ERR_NO_DEVICE) return 22;
  apex_bal_free(ds);
  printf("ok\n");
  return 0;
}
""")
    exe = tmp_path / "client"
    libdir = os.path.dirname(F.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lapex_gpu", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe), str(tmp_path / "tiny.txt")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_config_presets_match_reference():  # levenberg_marquardt.rs:319-359, 519-530
    lib = F.load_library()
    d, b = F.LmConfig(), F.LmConfig()
    lib.apex_lm_config_default(C.byref(d))
    lib.apex_lm_config_for_bundle_adjustment(C.byref(b))
    assert (d.max_iterations, d.cost_tolerance, d.parameter_tolerance, d.gradient_tolerance) == (50, 1e-6, 1e-8, 1e-10)
    assert (d.damping, d.damping_min, d.damping_max, d.damping_nu) == (1e-3, 1e-12, 1e12, 2.0)
    assert d.schur_variant == F.SCHUR_EXPLICIT and np.isnan(d.min_cost_threshold)
    assert (b.max_iterations, b.damping, b.schur_variant, b.schur_preconditioner) == (20, 1e-3, F.SCHUR_EXPLICIT_PCG, F.PRECOND_SCHUR_JACOBI)
    assert (b.cg_max_iterations, b.cg_tolerance) == (200, 1e-6)  # explicit_schur.rs:211-212
    # the oracle's presets are the same bytes
    o = F.LmConfig()
    oracle_lib().oracle_lm_config_for_bundle_adjustment(C.byref(o))
    assert bytes(o)[:24] == bytes(b)[:24] and o.cg_tolerance == b.cg_tolerance


@pytest.mark.skipif(HAVE_GPU, reason="checks the behaviour WITHOUT a CUDA device")
def test_no_device_means_error_not_fallback():
    with pytest.raises(F.ApexError) as e:
        GpuContext()
    assert e.value.status == F.ERR_NO_DEVICE


def test_block_cyclic_sharding_covers_and_balances():
    prob = synth.make_problem(20, 3000, 5.0, seed=2)
    cnt = np.bincount(prob.obs_pt, minlength=prob.npts)
    for nranks in (1, 2, 3, 8):
        parts = [shard_info(prob.obs_pt, prob.npts, nranks, r) for r in range(nranks)]
        block = parts[0][0]
        assert all(p[0] == block for p in parts) and block == 128
        owner = (np.arange(prob.npts) // block) % nranks
        assert sum(p[1] for p in parts) == prob.npts and sum(p[2] for p in parts) == prob.nobs
        for r, (_, npl, n) in enumerate(parts):
            assert npl == int((owner == r).sum()) and n == int(cnt[owner == r].sum())
            assert abs(n - prob.nobs / nranks) <= 0.2 * prob.nobs / nranks + cnt.max() * block, "balanced by observations"
    with pytest.raises(F.ApexError):
        shard_info(prob.obs_pt, prob.npts, 2, 2)


def test_layout_invariants_incl_long_tracks_and_unobserved_landmarks():
    """The static structure apex_problem_upload builds (host-only entry point): every observation in exactly one slot,
    landmarks never straddle a 256-slot chunk, a landmark with more than 256 observations gets its own chunks, and the
    camera-sorted lane order of every chunk is consistent - for every rank of a sharded problem."""
    prob = synth.make_problem(300, 3000, 6.0, seed=11)
    keep = (prob.obs_pt != 3) & (prob.obs_pt >= 2)        # landmarks 0,1 unobserved; landmark 3 rebuilt below
    extra = np.arange(prob.ncam, dtype=np.uint32)          # landmark 3 seen by all 300 cameras (> one chunk)
    p2 = BAProblem(camera_model=prob.camera_model, opt_flags=prob.opt_flags, pose=prob.pose, intr=prob.intr, pt=prob.pt,
                   obs_cam=np.concatenate([prob.obs_cam[keep], extra]), obs_pt=np.concatenate([prob.obs_pt[keep], np.full(prob.ncam, 3, np.uint32)]),
                   obs_uv=np.concatenate([prob.obs_uv[keep], np.zeros((prob.ncam, 2))]))
    s = layout_stats(p2)
    assert s.consistent == 1 and s.slots_used == s.nobs_local == p2.nobs
    assert s.nlong_tiles == 1 and s.nchunks == s.nnormal_chunks + 2      # 300 observations -> 2 chunks
    assert s.nobs_local / (s.nchunks * 256) > 0.9, "chunk fill"
    for nranks in (3, 4):     # the layout build owns landmarks by shifts and masks for powers of two, by division otherwise
        total = 0
        for r in range(nranks):
            sr = layout_stats(p2, nranks, r)
            assert sr.consistent == 1
            total += sr.nobs_local
            assert (sr.shard_block, sr.npts_local, sr.nobs_local) == shard_info(p2.obs_pt, p2.npts, nranks, r)
        assert total == p2.nobs
    bad = BAProblem(camera_model=prob.camera_model, opt_flags=prob.opt_flags, pose=prob.pose, intr=prob.intr, pt=prob.pt,
                    obs_cam=np.array([999], np.uint32), obs_pt=np.array([0], np.uint32), obs_uv=np.zeros((1, 2)))
    with pytest.raises(F.ApexError) as e:
        layout_stats(bad)
    assert e.value.status == F.ERR_INVALID_INPUT


def test_camera_sorted_lane_order_of_the_chunks():
    """Split slot order of the chunk kernel (host side): inside every normal chunk the camera-sorted lanes are a permutation of
    the point-major lanes, cameras ascend along them, warp-segments count the runs of one camera cut at warp boundaries and
    every warp-segment appears exactly once in its camera's list (layout self-check); a chunk of ~256 observations of
    ring-local tracks sees far fewer cameras than observations."""
    s = layout_stats(synth.make_problem(1500, 30000, 5.0, seed=6))
    assert s.consistent == 1 and s.nnormal_chunks > 400
    assert s.max_segments_per_chunk <= 256
    assert s.nnormal_chunks * 8 <= s.nsegments < 0.9 * s.nobs_local     # >= one segment per warp, fewer than observations
    tiny = layout_stats(synth.make_problem(3, 40, 2.5, seed=2))          # 3 cameras: every warp holds at most 3 runs
    assert tiny.consistent == 1 and tiny.nsegments <= 3 * 8 * tiny.nnormal_chunks


def test_generator_is_deterministic_and_bal_shaped():
    a, b = synth.make_shape("ladybug49"), synth.make_shape("ladybug49")
    for x, y in ((a.pose, b.pose), (a.pt, b.pt), (a.obs_uv, b.obs_uv), (a.obs_cam, b.obs_cam)):
        assert np.array_equal(x, y)
    assert (a.ncam, a.npts) == (49, 7776) and 20000 < a.nobs < 40000
    key = a.obs_pt.astype(np.int64) * a.ncam + a.obs_cam
    assert len(np.unique(key)) == a.nobs, "one observation per (camera, landmark)"
    assert a.pose_fixed[0] == 0x3F and not a.pose_fixed[1:].any()  # bin/bundle_adjustment.rs:294-298
    assert np.allclose(np.linalg.norm(a.pose[:, 3:], axis=1), 1.0)
    # every observation is in front of its camera at the initial values (BAL looks down -z)
    R = synth.quat_to_matrix(a.meta["truth_pose"][:, 3:])
    pc = np.einsum("nij,nj->ni", R[a.obs_cam], a.meta["truth_pt"][a.obs_pt]) + a.meta["truth_pose"][a.obs_cam, :3]
    assert (pc[:, 2] < 0).all()


@pytest.mark.parametrize("variant", [F.SCHUR_EXPLICIT, F.SCHUR_IMPLICIT, F.SCHUR_EXPLICIT_PCG], ids=["explicit", "implicit", "explicit_pcg"])
def test_oracle_lm_converges_and_variants_agree(variant):
    """Self-consistency of the restated oracle (SURVEY §8c iv): every Schur variant reduces the cost of a Ladybug-shaped
    problem and they reach the same neighbourhood."""
    prob = synth.make_problem(16, 600, 4.0, seed=7, self_calibration=False)
    o = OracleContext().upload(prob)
    cfg = o.default_config(True)
    cfg.schur_variant = variant
    cfg.max_iterations = 12
    res, trace = o.lm_solve(cfg)
    assert res.final_cost < 0.15 * res.initial_cost  # 2 % gross outliers keep a Huber floor
    assert res.iterations == len(trace) and res.successful_steps + res.unsuccessful_steps == res.iterations
    assert res.cost_evaluations == res.iterations + 1 and res.jacobian_evaluations == res.iterations
    for a, b in zip(trace, trace[1:]):
        assert b.cost <= a.cost + 1e-12, "accepted costs never increase"
    ref = OracleContext().upload(prob)
    cfg2 = ref.default_config(True)
    cfg2.schur_variant = F.SCHUR_EXPLICIT
    cfg2.max_iterations = 12
    r2, _ = ref.lm_solve(cfg2)
    assert abs(res.final_cost - r2.final_cost) <= 0.15 * r2.final_cost  # truncated PCG (200 its) lags the direct solve


def test_oracle_observer_feed():
    """OptObserver feed of the LM loop (parity_helpers.check_observer_feed) on the oracle; the GPU library runs the same check in -m gpu."""
    check_observer_feed(OracleContext, synth.make_problem(10, 300, 4.0, seed=11, self_calibration=True))


def test_oracle_step_application_and_parameter_norm_kats():
    """src/optimizer/mod.rs:1123-1180 restated on BA variables: apply_parameter_step advances an Rn variable by the step (x = 0, step 3
    -> 3) and apply_negative_parameter_step reverts it exactly on Rn (5 -> 7 -> 5); on SE3 the step goes through (+) (pose o Exp(step))
    and the negative step returns to the start only to first order (the reference's revert, mod.rs:343-356); compute_parameter_norm is the
    2-norm over the variables' storage (sqrt(3^2 + 4^2) = 5): 7 numbers per pose incl. the quaternion, intrinsics, landmarks; fixed
    indices are zeroed at the update only (problem.rs:185-289)."""
    prob = synth.make_problem(3, 12, 3.0, seed=5, self_calibration=True)
    prob.pt[:] = 0.0
    prob.pt[0] = [5.0, 0.0, 0.0]
    prob.pt_fixed = np.zeros(prob.npts, dtype=np.uint8); prob.pt_fixed[1] = 0b010      # y of landmark 1 is fixed
    o = OracleContext().upload(prob)
    pose0, intr0, pt0 = [a.copy() for a in o.params_download()]
    sc = np.zeros((prob.ncam, prob.dc)); sp = np.zeros((prob.npts, 3))
    sp[0] = [2.0, 0.0, 0.0]; sp[1] = [1.0, 1.0, 1.0]; sp[2] = [3.0, 0.0, 0.0]
    o.apply_step(sc, sp, +1.0)
    _, _, pt1 = o.params_download()
    assert pt1[0, 0] == 7.0 and pt1[2, 0] == 3.0 and list(pt1[1]) == [1.0, 0.0, 1.0]
    o.apply_step(sc, sp, -1.0)
    assert np.array_equal(o.params_download()[2], pt0), "Rn: the negative step is an exact revert"
    sc[1, :6] = [0.01, -0.02, 0.03, 0.004, -0.005, 0.006]
    sc[1, 6:] = 0.5
    o.apply_step(sc, np.zeros_like(sp), +1.0)
    pose1, intr1, _ = o.params_download()
    assert np.array_equal(pose1[[0, 2]], pose0[[0, 2]]) and not np.array_equal(pose1[1], pose0[1])
    assert np.allclose(intr1[1], intr0[1] + 0.5, rtol=0, atol=1e-12) and abs(np.linalg.norm(pose1[1, 3:]) - 1.0) < 1e-15
    o.apply_step(sc, np.zeros_like(sp), -1.0)
    pose2, intr2, _ = o.params_download()
    assert np.allclose(pose2[1], pose0[1], atol=1e-3) and not np.array_equal(pose2[1], pose0[1]), "SE3: (x (+) d) (+) (-d) = x only to first order"
    assert np.allclose(intr2[1], intr0[1], rtol=0, atol=1e-12)
    # parameter norm: one LM iteration whose step is (almost) zero - damping 1e30 - reports the norm of the variables' storage
    o = OracleContext().upload(prob)
    cfg = o.default_config(True)
    cfg.schur_variant = F.SCHUR_EXPLICIT
    cfg.max_iterations = 0
    cfg.damping, cfg.damping_max = 1e30, 1e32
    _, tr = o.lm_solve(cfg)
    pose, intr, pt = o.params_download()
    want = math.sqrt((pose ** 2).sum() + (intr ** 2).sum() + (pt ** 2).sum())
    assert abs(tr[0].parameter_norm - want) <= 1e-12 * want and tr[0].step_norm < 1e-20


def test_oracle_linear_solve_matches_scipy():
    """SURVEY §8c iii: the oracle's explicit Schur step solves (J^T J + lambda I) delta = -J^T r (checked densely)."""
    import scipy.linalg
    prob = synth.make_problem(6, 60, 4.0, seed=3)
    o = OracleContext().upload(prob)
    lam = 1e-2
    o.linearize(lam)
    r, jc, jp = o.get_linearization()
    dc, ncam, npts = prob.dc, prob.ncam, prob.npts
    n = ncam * dc + 3 * npts
    J = np.zeros((2 * prob.nobs, n))
    for i in range(prob.nobs):
        c, p = int(prob.obs_cam[i]), int(prob.obs_pt[i])
        J[2 * i:2 * i + 2, c * dc:(c + 1) * dc] = jc[i]
        J[2 * i:2 * i + 2, ncam * dc + 3 * p:ncam * dc + 3 * p + 3] = jp[i]
    H = J.T @ J + lam * np.eye(n)
    g = J.T @ r.reshape(-1)
    delta = scipy.linalg.cho_solve(scipy.linalg.cho_factor(H), -g)
    sc, sp, gnorm, _ = o.solve_augmented(F.SCHUR_EXPLICIT, lam)
    assert abs(gnorm - np.linalg.norm(g)) <= 1e-10 * gnorm
    assert np.abs(sc.reshape(-1) - delta[:ncam * dc]).max() <= 1e-6 * np.abs(delta).max()
    assert np.abs(sp.reshape(-1) - delta[ncam * dc:]).max() <= 1e-6 * np.abs(delta).max()
    # and the matrix-free operator is the Schur complement of H
    x = np.random.default_rng(0).standard_normal(ncam * dc)
    Hcc, Hcp, Hpp = H[:ncam * dc, :ncam * dc], H[:ncam * dc, ncam * dc:], H[ncam * dc:, ncam * dc:]
    S = Hcc - Hcp @ np.linalg.solve(Hpp, Hcp.T)
    assert np.abs(o.schur_matvec(x) - S @ x).max() <= 1e-9 * np.abs(S @ x).max()


def mixed_loss_problem(prob, seed=3):
    """Per-block loss functions (src/core/residual_block.rs:97-123): every observation draws one of four LossFunction instances."""
    table = [(F.LOSS_HUBER, 1.0), (F.LOSS_CAUCHY, 2.0), (F.LOSS_NONE,), (F.LOSS_TUKEY, 4.0)]
    idx = np.random.default_rng(seed).integers(0, len(table), prob.nobs).astype(np.uint8)
    return dataclasses.replace(prob, obs_loss=idx, loss_table=table, meta={}), table, idx


def test_oracle_per_block_loss_is_the_sum_over_loss_classes():
    """Oracle with a per-block loss table: the cost is the sum of the costs of the sub-problems that hold the observations of
    one loss each (uniform-loss path), a table whose entries are all the same loss reproduces the uniform problem bit for bit,
    and indices outside the table are rejected."""
    base = synth.make_problem(10, 300, 4.0, seed=17)
    prob, table, idx = mixed_loss_problem(base)
    o = OracleContext().upload(prob)
    total = 0.0
    for k, spec in enumerate(table):
        m = idx == k
        lp = tuple(spec[1:]) + (0.0,) * 4
        sub = dataclasses.replace(base, obs_cam=base.obs_cam[m], obs_pt=base.obs_pt[m], obs_uv=base.obs_uv[m], loss_id=spec[0], loss_params=lp[:4], meta={})
        total += OracleContext().upload(sub).cost()
    assert abs(o.cost() - total) <= 1e-12 * total
    same = dataclasses.replace(base, obs_loss=np.zeros(base.nobs, np.uint8), loss_table=[(base.loss_id,) + tuple(base.loss_params)], meta={})
    a, b = OracleContext().upload(same), OracleContext().upload(base)
    assert a.cost() == b.cost()
    a.linearize(1e-3); b.linearize(1e-3)
    assert all(np.array_equal(x, y) for x, y in zip(a.get_blocks(), b.get_blocks()))
    bad = dataclasses.replace(base, obs_loss=np.full(base.nobs, 7, np.uint8), loss_table=table, meta={})
    with pytest.raises(F.ApexError) as e:
        OracleContext().upload(bad)
    assert e.value.status == F.ERR_INVALID_INPUT
    with pytest.raises(F.ApexError) as e:
        layout_stats(bad)
    assert e.value.status == F.ERR_INVALID_INPUT


def test_oracle_jacobi_scaling_matches_a_dense_restatement():
    """process_jacobian_generic (src/optimizer/mod.rs:749-763) + compute_step_generic (levenberg_marquardt.rs:730-766), restated
    densely with numpy on a tiny problem: scaling = 1 / (1 + column norm), (Js^T Js + lambda I) dxs = -Js^T r, step = dxs * scaling,
    reported gradient norm = ||Js^T r||. The oracle's first LM iteration with use_jacobi_scaling must give that step."""
    prob = synth.make_problem(4, 30, 3.0, seed=5, self_calibration=True)
    o = OracleContext().upload(prob)
    lam = 1e-3
    o.linearize(lam)
    r, jc, jp = o.get_linearization()
    dc, ncd = prob.dc, prob.ncam * prob.dc
    n = ncd + 3 * prob.npts
    J = np.zeros((2 * prob.nobs, n))
    for i in range(prob.nobs):
        c, p = int(prob.obs_cam[i]), int(prob.obs_pt[i])
        J[2 * i:2 * i + 2, c * dc:(c + 1) * dc] = jc[i].reshape(2, dc)
        J[2 * i:2 * i + 2, ncd + 3 * p:ncd + 3 * p + 3] = jp[i].reshape(2, 3)
    sc = 1.0 / (1.0 + np.linalg.norm(J, axis=0))
    Js = J * sc[None, :]
    g = Js.T @ r.reshape(-1)
    dxs = np.linalg.solve(Js.T @ Js + lam * np.eye(n), -g)
    step = dxs * sc
    cfg = o.default_config(True)
    cfg.schur_variant = F.SCHUR_EXPLICIT
    cfg.max_iterations = 0
    cfg.damping = lam
    cfg.use_jacobi_scaling = 1
    res, tr = o.lm_solve(cfg)
    dca, dpa = o.get_step()
    assert np.abs(dca.reshape(-1) - step[:ncd]).max() <= 1e-8 * np.abs(step[:ncd]).max()
    assert np.abs(dpa.reshape(-1) - step[ncd:]).max() <= 1e-8 * np.abs(step[ncd:]).max()
    assert abs(tr[0].gradient_norm - np.linalg.norm(g)) <= 1e-10 * np.linalg.norm(g)
    assert abs(tr[0].step_norm - np.linalg.norm(step)) <= 1e-8 * np.linalg.norm(step)
    predicted = 0.5 * step @ (lam * step - g)     # compute_predicted_reduction with the UNSCALED step and the SCALED gradient
    assert abs(tr[0].predicted_reduction - predicted) <= 1e-8 * abs(predicted)
    cfg.use_jacobi_scaling = 0                      # and without it the oracle takes another step
    o2 = OracleContext().upload(prob)
    o2.lm_solve(cfg)
    assert np.abs(o2.get_step()[0] - dca).max() > 1e-6 * np.abs(dca).max()


def reference_calibration_problem(model):
    """The reference's calibration test inputs, value for value: truth from synth.make_calibration_scene (wall for Kannala-Brandt,
    generate_scene_points for double sphere), initial values from the reference's own index-driven Box-Muller noise
    (tests/camera_test_utils.rs:53-150: landmarks 1 cm seed 100, poses 2 cm / 1 deg seed 200 + 10 i, intrinsics 2 % seed 300;
    pose noise applied with SE3 right-plus)."""
    def ref_normal(std, index):
        u1 = ((index * 12345 + 67890) % 10000) / 10000.0
        u2 = ((index * 54321 + 98765) % 10000) / 10000.0
        return std * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    prob = synth.make_calibration_scene(model, scene="wall" if model == F.CAM_KANNALA_BRANDT else "hemisphere")
    truth_pt, truth_pose, truth_intr = prob.meta["truth_pt"], prob.meta["truth_pose"], prob.meta["truth_intr"]
    pt = truth_pt + np.array([[ref_normal(0.01, 100 + 3 * i + k) for k in range(3)] for i in range(truth_pt.shape[0])])
    lib = oracle_lib()
    pose = np.empty_like(truth_pose)
    for i in range(truth_pose.shape[0]):
        base = 200 + 10 * i
        tau = np.array([ref_normal(0.02, base + k) for k in range(3)] + [ref_normal(np.deg2rad(1.0), base + 3 + k) for k in range(3)])
        lib.oracle_se3_plus(F.ptr(np.ascontiguousarray(truth_pose[i])), F.ptr(tau), pose[i].ctypes.data)
    intr0 = np.array([truth_intr[i] * (1.0 + ref_normal(0.02, 300 + i)) for i in range(len(truth_intr))])
    return dataclasses.replace(prob, pose=pose, pt=pt, intr=np.tile(intr0, (prob.ncam, 1)), meta=dict(prob.meta))


@pytest.mark.parametrize("model,rel_tol,floor", [(F.CAM_KANNALA_BRANDT, [0.05, 0.05, 0.05, 0.05, 0.10, 0.20], 0.01), (F.CAM_DOUBLE_SPHERE, [0.05] * 4 + [0.10, 0.10], 0.1)],
                         ids=["kannala_brandt", "double_sphere"])
def test_oracle_calibration_scene_meets_the_reference_tests_criteria(model, rel_tol, floor):
    """The reference's own acceptance criteria for its multi-observation / shared-intrinsics graphs, on its own inputs
    (tests/camera_kannala_brandt_integration.rs:45-330, camera_double_sphere_integration.rs:38-310): LM (100 iterations, cost /
    parameter / gradient tolerances 1e-8 / 1e-8 / 1e-10, damping 1e-3) on 5 cameras x 200 points with ONE intrinsics variable, every
    pose_0 DOF fixed, must end in a converged status, reduce the cost by more than 85 %, reach a reprojection RMSE below 2 px and
    recover the intrinsics within the tests' per-parameter tolerances."""
    prob = reference_calibration_problem(model)
    o = OracleContext().upload(prob)
    cfg = o.default_config(False)
    cfg.schur_variant = F.SCHUR_EXPLICIT
    cfg.max_iterations, cfg.cost_tolerance, cfg.parameter_tolerance, cfg.gradient_tolerance, cfg.damping = 100, 1e-8, 1e-8, 1e-10, 1e-3
    res, tr = o.lm_solve(cfg)
    assert res.status in (0, 2, 3, 4)   # Converged | CostToleranceReached | ParameterToleranceReached | GradientToleranceReached (include/apex_gpu.h)
    assert (res.initial_cost - res.final_cost) / res.initial_cost > 0.85
    assert np.sqrt(res.final_cost / prob.nobs) < 2.0
    pose, intr, pt = o.params_download()
    assert np.array_equal(intr, np.tile(intr[0], (prob.ncam, 1))), "one intrinsics variable"
    truth = prob.meta["truth_intr"]
    for i, tol in enumerate(rel_tol):
        assert abs(intr[0, i] - truth[i]) / max(abs(truth[i]), floor) < tol, (i, intr[0, i], truth[i])
    for i in range(len(rel_tol), len(truth)):
        assert abs(intr[0, i] - truth[i]) < 0.25
    # the shared variable is not the per-camera problem: another trajectory
    per_cam = dataclasses.replace(prob, opt_flags=prob.opt_flags & ~F.OPT_SHARED_INTRINSICS, meta={})
    r2, _ = OracleContext().upload(per_cam).lm_solve(cfg)
    assert abs(r2.final_cost - res.final_cost) > 1e-9 * max(res.final_cost, 1e-12) or r2.iterations != res.iterations


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the contract's keys (runs the oracle on a bounded sample)."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--scale", "0.05"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "LM iterations/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_teacher_forced_parity_harness_on_cpu():
    """The harness the `-m gpu` suite uses (tests/parity_helpers.py), exercised here with the oracle in another summation
    order (reversed block sums, landmark sweeps spread over 8 threads) as the device under test: cost / gradient / backward
    agreement of the steps / back-substitution / trial cost / rho / damping / accept flag / parameters at every iterate of
    the oracle's trajectory."""
    from oracle_backend import OracleVariant
    from parity_helpers import teacher_forced_parity
    prob = synth.make_problem(16, 600, 4.0, seed=7, self_calibration=True)
    for variant, bw in ((F.SCHUR_IMPLICIT, 1e-9), (F.SCHUR_EXPLICIT, 1e-11)):
        rows = teacher_forced_parity(prob, variant, OracleVariant(reverse=True, parallel=True, threads=8).upload(prob), OracleContext().upload(prob),
                                     OracleContext().upload(prob), oracle_lib(), twin=OracleVariant(reverse=True).upload(prob), n_it=3)
        assert len(rows) == 3 and all(r["backward"] < bw for r in rows)


# ---- N > 1: observation/landmark sharding with replicated camera blocks, over gloo, world_size 2 ---------------
GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from apex_solver_b200 import synth
from apex_solver_b200.context import shard_info
from oracle_backend import OracleContext
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
prob = synth.make_problem(10, 400, 4.0, seed=5)
o = OracleContext().upload(prob)
lam = 1e-3
o.linearize(lam)
n = prob.ncam * prob.dc
x = np.random.default_rng(1).standard_normal(n)
block, npl, nloc = shard_info(prob.obs_pt, prob.npts, world, rank)
# what one rank of the GPU path computes: the H_cc term once (rank 0) + its landmark blocks' part of -H_cp Hpp^-1 H_cp^T x
ysum = np.zeros(n)
first = True
for b0 in range(rank * block, prob.npts, world * block):      # block-cyclic ownership
    ysum += o.schur_matvec_partial(x, b0, min(b0 + block, prob.npts), first and rank == 0)
    first = False
if first and rank == 0:
    ysum += o.schur_matvec_partial(x, 0, 0, True)
y = torch.from_numpy(ysum.copy())
dist.all_reduce(y)            # the single exchange step of the operator (ncam*dc doubles)
full = o.schur_matvec(x)
err = float(np.abs(y.numpy() - full).max() / np.abs(full).max())
cnt = torch.tensor([nloc], dtype=torch.int64); dist.all_reduce(cnt)
ok = err < 1e-12 and int(cnt.item()) == prob.nobs
print("RANK", rank, "err", err, "ok", ok, flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_sharded_operator_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-2000:]

"""Development experiment: what one of N ranks holds under different landmark ownership rules (contiguous, block-cyclic with several
block sizes), as a stand-alone single-rank problem: operator time, partial rows of the deterministic flush, PCG iteration time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("APEX_TAIL_TRACE", "1")
import numpy as np
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import BAProblem, GpuContext, layout_stats
full = synth.make_shape(sys.argv[2] if len(sys.argv) > 2 else "venice1778")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rules = [("contiguous", np.arange(full.npts) * N // full.npts == 0)] + [("block-cyclic-%d" % b, (np.arange(full.npts) // b) % N == 0) for b in (128, 1024, 8192, 32768)]
for name, owned in rules:
    keep = owned[full.obs_pt]
    new_id = np.cumsum(owned) - 1
    prob = BAProblem(camera_model=full.camera_model, opt_flags=full.opt_flags, pose=full.pose, intr=full.intr, pt=full.pt[owned],
                     obs_cam=full.obs_cam[keep], obs_pt=new_id[full.obs_pt[keep]].astype(np.uint32), obs_uv=full.obs_uv[keep],
                     loss_id=full.loss_id, loss_params=full.loss_params, pose_fixed=full.pose_fixed, intr_fixed=full.intr_fixed)
    st = layout_stats(prob)
    g = GpuContext().upload(prob)
    g.linearize(1e-3)
    ms = g.schur_matvec_bench(20, True)
    b = prob.nobs * 200 + prob.npts * 48 + prob.ncam * 792
    cfg = g.default_config(True); cfg.schur_variant = F.SCHUR_IMPLICIT; cfg.max_iterations = 3
    cfg.cost_tolerance = cfg.parameter_tolerance = cfg.gradient_tolerance = 0.0
    g.lm_solve(cfg)
    g.profile_read()
    res, _ = g.lm_solve(cfg)
    dev_ms = g.profile_read().lm_device_ms
    print(name, "nobs", prob.nobs, "cams touched", len(np.unique(prob.obs_cam)), "stand-alone operator ms %.4f" % ms, "GB/s %.0f" % (b / ms / 1e6),
          "| LM: %.2f ms per iteration, %d PCG iterations -> %.1f us per PCG iteration (incl. the once-per-LM-iteration kernels)" % (dev_ms / res.iterations, res.linear_iterations, 1e3 * dev_ms / max(res.linear_iterations, 1)),
          "| layout:", {k: getattr(st, k) for k in ("nnormal_chunks", "mv_ranges", "mv_nwindows", "mv_rows")}, flush=True)
    g.close()

"""Development experiment: operator time on 1/8 of the Venice landmarks, contiguous vs block-cyclic selection."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from apex_solver_b200 import synth
from apex_solver_b200.context import BAProblem, GpuContext
full = synth.make_shape("venice1778")
N = 8
for name, owned in (("contiguous", np.arange(full.npts) < full.npts // N), ("block-cyclic-128", (np.arange(full.npts) // 128) % N == 0)):
    keep = owned[full.obs_pt]
    new_id = np.cumsum(owned) - 1
    prob = BAProblem(camera_model=full.camera_model, opt_flags=full.opt_flags, pose=full.pose, intr=full.intr, pt=full.pt[owned],
                     obs_cam=full.obs_cam[keep], obs_pt=new_id[full.obs_pt[keep]].astype(np.uint32), obs_uv=full.obs_uv[keep],
                     loss_id=full.loss_id, loss_params=full.loss_params, pose_fixed=full.pose_fixed, intr_fixed=full.intr_fixed)
    g = GpuContext().upload(prob)
    g.linearize(1e-3)
    ms = g.schur_matvec_bench(20, True)
    dc = prob.dc
    b = prob.nobs * 200 + prob.npts * 48 + prob.ncam * 792
    print(name, "nobs", prob.nobs, "cams touched", len(np.unique(prob.obs_cam)), "matvec ms %.4f" % ms, "GB/s %.0f" % (b / ms / 1e6), flush=True)
    g.close()

"""Development aid: phases of the FIRST upload on a new context (APEX_LAYOUT_TIMING) + solve + download."""
import os, sys, time
os.environ["APEX_LAYOUT_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
prob = synth.make_shape("venice1778")
warm = GpuContext().upload(synth.make_problem(20, 500, 4.0))   # CUDA context, library load
for rep in range(2):
    g = GpuContext()
    t0 = time.perf_counter(); g.upload(prob); t1 = time.perf_counter()
    cfg = g.default_config(True); cfg.schur_variant = F.SCHUR_IMPLICIT; cfg.max_iterations = 9
    cfg.cost_tolerance = cfg.parameter_tolerance = cfg.gradient_tolerance = 0.0
    res, _ = g.lm_solve(cfg); t2 = time.perf_counter()
    out = g.params_download(); t3 = time.perf_counter()
    print(f"COLD context {rep}: upload {1e3*(t1-t0):.1f} ms  solve({res.iterations} it) {1e3*(t2-t1):.1f} ms  download {1e3*(t3-t2):.1f} ms", flush=True)
    g.close()

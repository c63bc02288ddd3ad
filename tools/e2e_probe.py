"""Development aid: where the end-to-end time of bench.py goes (second upload on a warm context, like the bench's e2e leg)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
prob = synth.make_shape("venice1778")
g = GpuContext().upload(prob)
cfg = g.default_config(True); cfg.schur_variant = F.SCHUR_IMPLICIT; cfg.max_iterations = 2
g.lm_solve(cfg)
os.environ["APEX_LAYOUT_TIMING"] = "1"
for _ in range(2):
    t0 = time.perf_counter(); g.upload(prob); t1 = time.perf_counter()
    res, _ = g.lm_solve(cfg); t2 = time.perf_counter()
    out = g.params_download(); t3 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.1f} ms  solve({res.iterations} it) {1e3*(t2-t1):.1f} ms  download {1e3*(t3-t2):.1f} ms", flush=True)

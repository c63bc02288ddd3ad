import os, sys, subprocess, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
from apex_solver_b200.bal import dataset_from_problem, load_bal
path = "/tmp/problem-12-400-pre.txt"
prob0 = synth.make_problem(12, 400, 4.0, seed=41)
if not os.path.exists(path):
    dataset_from_problem(prob0).write(path)
prob = load_bal(path).problem(optimization_type="bundle-adjustment")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else F.SCHUR_IMPLICIT
g = GpuContext().upload(prob)
cfg = g.default_config(True); cfg.schur_variant = variant; cfg.max_iterations = 20
res, tr = g.lm_solve(cfg)
print("py", variant, repr(res.final_cost), res.iterations, res.linear_iterations, [repr(t.cost) for t in tr[:4]])
exe = os.path.join(os.path.dirname(F.LIB_PATH), "bundle_adjustment")
flags = ["-s", "matrix-free"] if variant == F.SCHUR_IMPLICIT else []
r = subprocess.run([exe, path, "-t", "bundle-adjustment", "-v"] + flags, capture_output=True, text=True, timeout=300)
print("cli", re.search(r"Final cost: (\S+)", r.stdout).group(1), re.findall(r"iter\s+\d+\s+cost (\S+)", r.stdout)[:4], re.findall(r"linear iterations (\d+)", r.stdout)[:6])

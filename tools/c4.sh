#!/bin/bash
# config 4 at full scale: Kannala-Brandt, 2000 cams / 1M pts / ~6M obs, explicit Schur (n = 28000), 2 LM iterations + launch list
python - <<'PY' > gpurun_out/c4_full.log 2>&1
import sys, time, json
sys.path.insert(0, ".")
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
prob = synth.make_shape("kb2000")
g = GpuContext().upload(prob)
cfg = g.default_config(True); cfg.schur_variant = F.SCHUR_EXPLICIT; cfg.max_iterations = 1
cfg.cost_tolerance = cfg.parameter_tolerance = cfg.gradient_tolerance = 0.0
t = time.time(); res, tr = g.lm_solve(cfg); dt = time.time() - t
print(json.dumps({"ncam": prob.ncam, "npts": prob.npts, "nobs": prob.nobs, "dc": prob.dc, "iters": res.iterations, "lm_s": dt, "cost0": res.initial_cost, "cost1": res.final_cost,
                  "accepted": res.successful_steps, "iter_ms": [x.iter_time_ms for x in tr]}))
PY
tail -2 gpurun_out/c4_full.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4.csv python - <<'PY' > /dev/null 2>&1
import sys
sys.path.insert(0, ".")
from apex_solver_b200 import _ffi as F, synth
from apex_solver_b200.context import GpuContext
prob = synth.make_shape("kb2000", scale=0.5)
g = GpuContext().upload(prob)
cfg = g.default_config(True); cfg.schur_variant = F.SCHUR_EXPLICIT; cfg.max_iterations = 0
cfg.cost_tolerance = cfg.parameter_tolerance = cfg.gradient_tolerance = 0.0
g.lm_solve(cfg)
PY
ls -la gpurun_out/launches_c4.csv

#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "linearize_blocks or solve_augmented or long_tracks or unobserved or fallback or deterministic or lm_solve_parity" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log | cut -c1-300
for impl in chunk pp; do
APEX_MATVEC_IMPL=$impl timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 20 > gpurun_out/probe_$impl.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_$impl.log").read().strip().splitlines()[-1])
print("$impl", {k: round(d[k], 4) for k in d if k.startswith("matvec")})
PY
done

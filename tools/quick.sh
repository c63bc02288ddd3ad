#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -x -k "linearize_blocks or solve_augmented or long_tracks or unobserved" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-300
timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 20 > gpurun_out/probe_pp.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_pp.log").read().strip().splitlines()[-1])
print("venice", {k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("linearize_ms", "cost_ms")})
PY

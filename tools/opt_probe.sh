#!/bin/bash
# development aid: A/B of the chunk kernel's options (APEX_MV_OPT bit mask) on the Venice shape
for o in ${OPTS:-0 1 2 4 3 7}; do
  APEX_MV_WINDOW=0 APEX_MV_OPT=$o timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 30 > gpurun_out/probe_opt_$o.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_opt_$o.log").read().strip().splitlines()[-1])
    print("OPT=$o", {k: round(d[k], 4) for k in d if k.startswith("matvec")})
except Exception as e:
    print("OPT=$o failed", e); print(open("gpurun_out/probe_opt_$o.log").read()[-1500:])
PY
done

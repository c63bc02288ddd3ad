#!/bin/bash
# development aid: one ncu --set full capture of the persistent operator kernel at the Venice shape + timing probe
timeout 400 python tools/probe.py --shape venice1778 --iters 1 > gpurun_out/probe_venice1778.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_venice1778.log").read().strip().splitlines()[-1])
print({k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("linearize_ms", "cost_ms")})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_matvec_persist -s 2 -c 1 -o gpurun_out/prof_matvec_v2 -f python tools/probe.py --shape venice1778 --iters 1 --reps 3 > gpurun_out/ncu_matvec.log 2>&1
tail -2 gpurun_out/ncu_matvec.log

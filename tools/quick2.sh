#!/bin/bash
for d in 0 1; do
APEX_DEBUG_MATVEC=$d APEX_MATVEC_IMPL=tileseg timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 20 > gpurun_out/probe_x.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/probe_x.log").read().strip().splitlines()[-1])
print("tileseg debug $d", {k: round(d[k], 4) for k in d if k.startswith("matvec")})
PY
done
APEX_MATVEC_IMPL=tileseg timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_tile_kernel -s 3 -c 1 -o gpurun_out/prof_tileseg -f python tools/probe.py --shape venice1778 --iters 1 --reps 3 > gpurun_out/ncu_tileseg.log 2>&1; tail -1 gpurun_out/ncu_tileseg.log

#!/bin/bash
# development aid: reductions-suppressed bounds + ncu capture of the window kernel
for cfg in "0 8 0" "0 8 1" "320 8 0" "320 8 1" "320 16 1"; do
  set -- $cfg
  APEX_MV_OPT=3 APEX_MV_WINDOW=$1 APEX_MV_GROUP=$2 APEX_DEBUG_MATVEC=$3 timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 30 > gpurun_out/probe_dbg_$1_$2_$3.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_dbg_$1_$2_$3.log").read().strip().splitlines()[-1])
    print("W=$1 G=$2 debug=$3", {k: round(d[k], 4) for k in d if k.startswith("matvec_ms")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/probe_dbg_$1_$2_$3.log").read()[-1500:])
PY
done
APEX_MV_WINDOW=320 APEX_MV_GROUP=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_window_kernel -s 2 -c 1 -o gpurun_out/prof_window -f python tools/probe.py --shape venice1778 --iters 1 --reps 3 > gpurun_out/ncu_window.log 2>&1
tail -2 gpurun_out/ncu_window.log

#!/bin/bash
# development aid: chunk kernel vs window kernel on the Venice shape + GPU tests
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in ${CFGS:-"0 8" "320 8"}; do
  set -- $cfg
  APEX_MV_WINDOW=$1 APEX_MV_GROUP=$2 timeout 400 python tools/probe.py --shape venice1778 --iters 2 --reps 20 > gpurun_out/probe_win_$1_$2.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_win_$1_$2.log").read().strip().splitlines()[-1])
    print("W=$1 G=$2", {k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("lm_it_per_s", "cost1", "pcg_iters", "upload_s")})
except Exception as e:
    print("W=$1 G=$2 failed", e); print(open("gpurun_out/probe_win_$1_$2.log").read()[-1500:])
PY
done

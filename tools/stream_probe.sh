#!/bin/bash
# development aid: stream kernel vs chunk kernel on the Venice shape + GPU tests (bounded by timeouts)
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "stream or window or fallback or deterministic or matvec or solve_augmented or long_tracks" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_quick.log | cut -c1-300
for cfg in "1 0" "1 1"; do
  set -- $cfg
  APEX_MV_WINDOW=0 APEX_MV_STREAM=$1 APEX_DEBUG_MATVEC=$2 timeout 300 python tools/probe.py --shape venice1778 --iters 2 --reps 30 > gpurun_out/probe_stream_$1_$2.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_stream_$1_$2.log").read().strip().splitlines()[-1])
    print("stream=$1 debug=$2", {k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("lm_it_per_s", "cost1", "pcg_iters")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/probe_stream_$1_$2.log").read()[-1500:])
PY
done

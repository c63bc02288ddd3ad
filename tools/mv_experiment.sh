#!/bin/bash
# development aid: matvec timing under the debug probes + one ncu capture of the tile kernel
for d in 0 1 2; do
  APEX_DEBUG_MATVEC=$d timeout 300 python tools/probe.py --shape venice1778 --iters 1 > gpurun_out/probe_dbg$d.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/probe_dbg$d.log").read().strip().splitlines()[-1])
print("debug $d", {k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("linearize_ms", "cost_ms")})
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:schur_tile_kernel -s 2 -c 2 -o gpurun_out/prof_matvec_v1 -f python tools/probe.py --shape venice1778 --iters 1 --reps 3 > gpurun_out/ncu_matvec.log 2>&1
tail -3 gpurun_out/ncu_matvec.log
ls -la gpurun_out/

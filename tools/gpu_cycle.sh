#!/bin/bash
# development aid: GPU test suite + probes, logs into gpurun_out/
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200
for shape in trafalgar257 venice1778; do
  timeout 400 python tools/probe.py --shape $shape --iters 3 > gpurun_out/probe_$shape.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_$shape.log").read().strip().splitlines()[-1])
    print("$shape", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k not in ("trace",)})
    print("   trace", d["trace"])
except Exception as e:
    print("$shape probe failed", e); print(open("gpurun_out/probe_$shape.log").read()[-1500:])
PY
done

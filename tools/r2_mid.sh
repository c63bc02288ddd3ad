#!/bin/bash
# mid-round check on one GPU: smoke, the whole GPU suite, the default bench line (no reference arm)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_n1_mid.json 2> gpurun_out/bench_n1_mid.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n1_mid.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["warm_value"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step")}, "launches", d["gpu_launches"], d["clocks"], "pcg", d["config"]["pcg_iterations"], d["config"]["final_cost"])
PY

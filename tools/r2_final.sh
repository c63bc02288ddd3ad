#!/bin/bash
# round-end check on one GPU: smoke, the whole GPU suite, the default bench line + the reference arm
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["warm_value"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step")}, "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json

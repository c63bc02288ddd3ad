#!/bin/bash
# full single-GPU cycle: GPU test suite, bench line (with the CPU leg), reference arm, launch list, one ncu --set full capture, upload timing
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["warm_value"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step")}, "cpu", d["cpu_baseline"]["value"])
PY
APEX_LAYOUT_TIMING=1 timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; tail -40 gpurun_out/e2e_probe.log
bash tools/r2_launches.sh A=0 | head -14
bash tools/r2_ncu.sh

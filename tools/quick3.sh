#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "mv_ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "pcg", d["config"]["pcg_iterations"], "final", d["config"]["final_cost"], "launches", d["gpu_launches"])
PY
timeout 300 python tools/configs.py --only "C2 trafalgar257 implicit" 2>&1 | tail -1 | cut -c1-500

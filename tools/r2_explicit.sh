#!/bin/bash
# explicit-Schur cycle: the explicit-path tests, then the C4 Kannala-Brandt line (formation of S / Cholesky split in the roofline block)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "explicit or c4 or shared_intrinsics or solve_augmented or ladybug or bal_file or observer" > gpurun_out/pytest_explicit.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_explicit.log | cut -c1-300
for shape in ${@:-kb2000}; do
timeout 900 python bench.py --shape $shape --variant explicit --steps 3 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_${shape}_explicit.json 2> gpurun_out/e.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${shape}_explicit.json").read().strip().splitlines()[-1])
print("$shape value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"], "final", d["config"].get("final_cost"), d["config"].get("accepted_steps"))
PY
done

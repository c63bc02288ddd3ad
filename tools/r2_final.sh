#!/bin/bash
# round-end check on one GPU: smoke, the whole GPU suite, the default bench line + the reference arm, then the ncu passes behind profiles/
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
if [ "$1" = "ncu" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_chunk_kernel -s 8 -c 1 -o gpurun_out/prof_matvec_cur -f python bench.py --steps 1 --warmup 0 --cpu-baseline 0 > gpurun_out/ncu_full.log 2>&1; echo "operator capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_syrk128 -s 4 -c 1 -o gpurun_out/prof_syrk128 -f python -c "
import sys; sys.path.insert(0, '.')
from apex_solver_b200.context import GpuContext
g = GpuContext(); print(g.dense_cholesky_bench(14016, 1))
" > gpurun_out/ncu_syrk.log 2>&1; echo "syrk128 capture rc=$?"
ls -la gpurun_out/*.ncu-rep
fi

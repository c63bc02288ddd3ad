#!/bin/bash
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python tools/probe.py --shape venice1778 --iters 20 --reps 3 > gpurun_out/probe_venice20.log 2>&1; python - <<PY
import json
d = json.loads(open("gpurun_out/probe_venice20.log").read().strip().splitlines()[-1])
print("venice 20 iters:", d["lm_iters"], d["lm_s"], d["status"], d["cost0"], d["cost1"]); print(d["trace"])
PY

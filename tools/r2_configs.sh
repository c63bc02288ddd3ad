#!/bin/bash
# bench lines of the other BASELINE.json configs on one GPU: C4 (KB / DS, explicit Schur + dense Cholesky), C5 at N=1, C1, C2
run() { name=$1; shift; timeout 1500 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$name value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3), "accepted", d["config"]["accepted_steps"], "final", d["config"]["final_cost"], "init", d["config"]["initial_cost"],
          "roofline", {k: r.get(k) for k in ("bound", "achieved", "frac", "avg_launch_ms", "share_of_step", "n", "schur_form_ms", "schur_form_share_of_step")})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
run kb2000_explicit --shape kb2000 --variant explicit --steps 3 --warmup 1 --cpu-baseline 0
run ds2000_explicit --shape ds2000 --variant explicit --steps 3 --warmup 1 --cpu-baseline 0
run final13682_n1 --shape final13682 --steps 8 --warmup 1 --cpu-baseline 0
run ladybug49_explicit --shape ladybug49 --variant explicit --steps 10 --warmup 3 --cpu-baseline 0
run trafalgar257 --shape trafalgar257 --steps 10 --warmup 3 --cpu-baseline 0

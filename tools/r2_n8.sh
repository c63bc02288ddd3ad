#!/bin/bash
# 8-GPU lines: Venice shape (the metric), with / without the fused PCG tail, and the Final-13682 shape (configs[4])
N=${1:-8}
bash tools/scale2.sh $N APEX_PCG_TAIL=0 2>&1 | cut -c1-330
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --shape final13682 --steps 8 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_final13682_n$N.json 2> gpurun_out/bench_final13682_n$N.err; echo "final rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_final13682_n$N.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("final13682 N=$N value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3), "accepted", d["config"]["accepted_steps"], "final", d["config"]["final_cost"], "init", d["config"]["initial_cost"], "pcg", d["config"]["pcg_iterations"],
          "roofline", {k: r.get(k) for k in ("achieved", "frac", "avg_launch_ms", "share_of_step")}, "parity", {k: d["config"]["parity_vs_n1"][k] for k in ("accept_pattern_equal", "pcg_iterations_equal", "max_rel_trial_cost_diff", "max_backward_step_agreement")} if d["config"].get("parity_vs_n1") else None)
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_final13682_n$N.err").read()[-2500:])
PY

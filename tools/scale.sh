#!/bin/bash
# development aid: bench at N ranks (run under gpurun --gpus N)
N=$1
nvidia-smi -L | head -$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "mv_ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "pcg", d["config"]["pcg_iterations"], "final", d["config"]["final_cost"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n$N.err").read()[-2000:])
PY

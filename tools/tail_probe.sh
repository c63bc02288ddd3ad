#!/bin/bash
# development aid: fused PCG tail vs three-kernel path (tests, then LM it/s on the Venice shape at N ranks)
N=${1:-1}
echo skip-pytest

for t in 16 0; do
  if [ "$N" = "1" ]; then
    APEX_PCG_TAIL=$t timeout 600 python bench.py --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_tail_${t}_n$N.json 2> gpurun_out/bench_tail_${t}_n$N.err
  else
    APEX_PCG_TAIL=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + t)) bench.py --gpus $N --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_tail_${t}_n$N.json 2> gpurun_out/bench_tail_${t}_n$N.err
  fi
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_tail_${t}_n$N.json") if l.startswith("{")][-1]
    print("tail=$t N=$N value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "mv_ms", round(d["roofline"]["avg_launch_ms"], 4), "pcg", d["config"]["pcg_iterations"], "final", d["config"]["final_cost"], "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_tail_${t}_n$N.err").read()[-1500:])
PY
done

#!/bin/bash
# development aid: bench at N ranks for a few environment settings (run under gpurun --gpus N): tools/scale2.sh N "ENV1=a ENV2=b" "ENV3=c" ...
N=$1; shift
run() {
  tag=$(echo "$*" | tr ' =' '__'); [ -z "$tag" ] && tag=default
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; rc=$?
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N [$*] rc=$rc value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "warm", round(d["e2e"].get("warm_value", 0), 2), "frac", round(d["roofline"]["frac"], 3), "mv_ms", round(d["roofline"]["avg_launch_ms"], 4), "share", round(d["roofline"]["share_of_step"], 3), "pcg", d["config"]["pcg_iterations"], "final", d["config"]["final_cost"], "parity", d["config"].get("parity_vs_n1"))
except Exception as e:
    print("N=$N [$*] rc=$rc failed", e); print(open("gpurun_out/bench_n${N}_$tag.err").read()[-1500:])
PY
}
run A=0
for e in "$@"; do run $e; done

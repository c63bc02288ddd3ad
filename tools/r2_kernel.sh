#!/bin/bash
# round-2 operator-kernel cycle: GPU test suite, timing probe (L2 flushed / not), bench line, one ncu --set full capture
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python tools/probe.py --shape venice1778 --iters 3 > gpurun_out/probe_venice1778.log 2>&1
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/probe_venice1778.log").read().strip().splitlines()[-1])
    print({k: round(d[k], 4) for k in d if k.startswith("matvec") or k in ("linearize_ms", "cost_ms", "lm_it_per_s", "upload_s")})
except Exception as e:
    print("probe failed", e); print(open("gpurun_out/probe_venice1778.log").read()[-1500:])
PY
APEX_DEBUG_MATVEC=1 timeout 400 python tools/probe.py --shape venice1778 --iters 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reductions suppressed:', {k: round(d[k],4) for k in d if k.startswith('matvec')})"
for det in 0 1; do APEX_DETERMINISTIC=$det timeout 400 python tools/probe.py --shape venice1778 --iters 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deterministic=$det:', {k: round(d[k],4) for k in d if k.startswith('matvec')})"; done
for st in 0 1; do for det in 0 1; do APEX_MV_STAGED=$st APEX_DETERMINISTIC=$det timeout 400 python tools/probe.py --shape venice1778 --iters 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('staged=$st det=$det:', {k: round(d[k],4) for k in d if k.startswith('matvec')})"; done; done
for w in 320 600; do APEX_MV_WINDOW=$w timeout 400 python tools/probe.py --shape venice1778 --iters 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('window=$w:', {k: round(d[k],4) for k in d if k.startswith('matvec')})"; done
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["warm_value"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_chunk_kernel -s 8 -c 1 -o gpurun_out/prof_matvec_cur -f python bench.py --steps 1 --warmup 0 --cpu-baseline 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200

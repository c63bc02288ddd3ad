#!/bin/bash
# PCG-tail cycle: the operator / solve unit tests, then LM timing + tail trace for a few switch settings on the full Venice shape
# and on an eighth of its landmarks (what one of eight ranks holds)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "fused_tail or solve_augmented or deterministic or pcg_iteration or windows_and_run or properties_trafalgar or teacher_forced and not c4 and not venice and not final" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_quick.log | cut -c1-300
probe() { frac=$1; shift; env "$@" APEX_TAIL_TRACE=1 timeout 400 python tools/probe.py --shape venice1778 --iters 6 --pts-frac $frac 2>gpurun_out/probe_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frac $frac $*:', {k: round(d[k],4) for k in d if k in ('matvec_ms_flush1','lm_it_per_s','pcg_iters','cost1')})"; grep "tail trace" gpurun_out/probe_err.log | tail -1 | cut -c1-300; grep -i "error\|Traceback" gpurun_out/probe_err.log | head -3; }
for f in 1.0 0.125; do
  probe $f A=0
  probe $f APEX_PDL=0
  for extra in "$@"; do probe $f $extra; done
done

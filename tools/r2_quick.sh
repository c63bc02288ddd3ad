#!/bin/bash
# quick operator-kernel cycle: the operator / solve unit tests, then the timing probe for a few switch settings
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "fused_tail or linearize_blocks or solve_augmented or deterministic or properties_trafalgar or teacher_forced and not c4 and not venice and not final" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_quick.log | cut -c1-300
probe() { env "$@" timeout 400 python tools/probe.py --shape venice1778 --iters 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*:', {k: round(d[k],4) for k in d if k.startswith('matvec') or k in ('lm_it_per_s','pcg_iters')})"; }
probe A=0
probe APEX_DETERMINISTIC=0
probe APEX_PCG_TAIL=0
for extra in "$@"; do probe $extra; done

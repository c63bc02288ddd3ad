#!/bin/bash
# operator variants on an eighth of the Venice landmarks (what one of eight ranks holds), single GPU
probe() { frac=$1; shift; env "$@" APEX_TAIL_TRACE=1 timeout 400 python tools/probe.py --shape venice1778 --iters 6 --pts-frac $frac 2>gpurun_out/probe_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frac $frac $*:', {k: round(d[k],4) for k in d if k in ('matvec_ms_flush1','matvec_ms_flush0','lm_it_per_s','pcg_iters','cost1')})"; grep "tail trace" gpurun_out/probe_err.log | tail -1 | cut -c100-330; grep -i "error\|Traceback" gpurun_out/probe_err.log | head -3; }
probe 0.125 A=0
for extra in "$@"; do probe 0.125 $extra; done

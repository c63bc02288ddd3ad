#!/bin/bash
APEX_MV_TIMING=1 timeout 400 python tools/probe.py --shape venice1778 --iters 1 --reps 2 > gpurun_out/probe_timing.log 2> gpurun_out/mv_timing.log
tail -4 gpurun_out/mv_timing.log

#!/bin/bash
# development aid: ncu --set full of the once-per-LM-iteration kernels (linearize, camera accumulation, cost, RHS, back-substitution) on the Venice shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linearize_tile_kernel|camera_accum_kernel|cost_tile_kernel|schur_chunk_kernel<9, 1>|schur_chunk_kernel<9, 2>' -c 6 -o gpurun_out/prof_lin -f python tools/probe.py --shape venice1778 --iters 1 --reps 1 > gpurun_out/ncu_lin.log 2>&1
tail -2 gpurun_out/ncu_lin.log

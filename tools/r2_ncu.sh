#!/bin/bash
# one ncu --set full capture of the operator kernel inside the bench workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_chunk_kernel -s 8 -c 1 -o gpurun_out/prof_matvec_cur -f python bench.py --steps 1 --warmup 0 --cpu-baseline 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200

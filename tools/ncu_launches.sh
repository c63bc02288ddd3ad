#!/bin/bash
# ncu launch list of the bench command (cold-cache, serialised: compare SHARES) + one full capture of the operator kernel
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schur_chunk_kernel -s 8 -c 1 -o gpurun_out/prof_matvec_cur -f python bench.py --steps 1 --warmup 0 --cpu-baseline 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | tail -5

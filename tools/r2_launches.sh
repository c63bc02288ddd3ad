#!/bin/bash
# ncu launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout 1200 env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_profiles.py r02 gpurun_out/launches.csv none | head -30

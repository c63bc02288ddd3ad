#!/bin/bash
timeout 600 ncu --set full --clock-control none -k regex:chol_syrk_kernel -s 3 -c 1 -o gpurun_out/prof_syrk -f python tools/configs.py --only "C4 kb2000 x0.5" > gpurun_out/ncu_syrk.log 2>&1; tail -1 gpurun_out/ncu_syrk.log | cut -c1-120

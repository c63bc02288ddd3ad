#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "explicit or solve_augmented or ladybug or trafalgar_full or long_tracks or fixed" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-300
timeout 600 python tools/configs.py --only "explicit" 2>&1 | grep -E "C3|C4|C2|C1 ladybug49 explicit\"" | cut -c1-420
timeout 600 ncu --set full --clock-control none -k regex:chol_syrk_kernel -s 60 -c 1 -o gpurun_out/prof_syrk -f python tools/configs.py --only "C4 kb2000 x0.5" > gpurun_out/ncu_syrk.log 2>&1; tail -1 gpurun_out/ncu_syrk.log | cut -c1-120

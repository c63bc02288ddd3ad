#!/bin/bash
# round-2 first GPU call: GPU test suite (with durations) + the default bench line
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=25 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err

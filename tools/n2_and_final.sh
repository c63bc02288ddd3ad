#!/bin/bash
# development aid: 2-GPU parity + bench, then (rank 0 only) the Final-13682 shape on one GPU
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_parity.py > gpurun_out/mgpu_parity_n2.log 2>&1; echo "parity rc=$?"; grep -E "N=|MGPU|Error|error" gpurun_out/mgpu_parity_n2.log | cut -c1-600 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-baseline 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.json | cut -c1-1500
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --shape final13682 --cpu-baseline 0 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "final rc=$?"; tail -c 2500 gpurun_out/bench_final_n1.json; tail -3 gpurun_out/bench_final_n1.err

#!/bin/bash
# development aid: 2-GPU parity tool + bench
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_parity.py > gpurun_out/mgpu_parity_n2.log 2>&1; echo "parity rc=$?"; grep -E "N=|MGPU|Error|error" gpurun_out/mgpu_parity_n2.log | cut -c1-300 | tail -5
bash tools/scale.sh 2

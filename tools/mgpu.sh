#!/bin/bash
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_parity.py > gpurun_out/mgpu_parity_n$N.log 2>&1; echo "parity rc=$?"; grep -E "N=|MGPU|Error|error" gpurun_out/mgpu_parity_n$N.log | cut -c1-600 | tail -8
bash tools/scale.sh $N

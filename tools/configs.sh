#!/bin/bash
timeout 1500 python tools/configs.py > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log | cut -c1-700

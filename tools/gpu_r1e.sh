#!/bin/bash
# one gpurun call (1 GPU): tensor-core training kernels vs the SIMT kernels, then the GPU suite
mkdir -p gpurun_out
timeout 300 python tools/tc_train_check.py --time > gpurun_out/tc_check.log 2>&1
echo "tc_check rc=$?"
cat gpurun_out/tc_check.log | tail -25
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r1e.log
cat gpurun_out/tests_r1e.log | tail -15

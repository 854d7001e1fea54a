#!/bin/bash
# one gpurun call (1 GPU): tensor-core tests, full bench, ncu launch list + full capture of the hot kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -k "tensor_core or graph" -x -q 2>&1 | tail -25 > gpurun_out/tests_tc.log
cat gpurun_out/tests_tc.log | tail -12
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
tail -c 1500 gpurun_out/bench_r1b.err
cat gpurun_out/bench_r1b.json | cut -c1-4000
SMALL="python bench.py --steps 6 --warmup 3 --eval-users 32 --no-cpu-baseline"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1b.csv $SMALL > gpurun_out/ncu1.log 2>&1
timeout 300 $SMALL > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_row_scores_tc|k_row_scores_splitk|k_bpr_bwd|k_adam_all' -s 8 -c 8 -o gpurun_out/prof_r1b $SMALL > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-300

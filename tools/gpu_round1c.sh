#!/bin/bash
# one gpurun call (1 GPU): whole GPU suite, smoke, full bench, ncu launch list + full capture of the tensor-core kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/tests_r1c.log
cat gpurun_out/tests_r1c.log | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -c 800 gpurun_out/bench_r1c.err
cat gpurun_out/bench_r1c.json | cut -c1-6000
SMALL="python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1c.csv $SMALL > gpurun_out/ncu1.log 2>&1
timeout 300 $SMALL > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_row_scores_tc|k_rank_eval' -c 3 -o gpurun_out/prof_r1c_tc $SMALL > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200

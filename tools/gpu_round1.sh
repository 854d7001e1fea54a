#!/bin/bash
# one gpurun call: parity tests, bench, ncu launch list + full capture of the three hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests.log
python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.err
cat gpurun_out/bench.json
SMALL="python bench.py --steps 6 --warmup 3 --eval-users 32 --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu1.log 2>&1
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_row_scores|k_bpr_bwd|k_adam_sweep' -s 12 -c 6 -o gpurun_out/prof_r1 $SMALL > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
cat gpurun_out/tests.log

#!/bin/bash
# one gpurun call (1 GPU): GPU suite, smoke, full bench (ours + reference arm), ncu launch list + full capture of the
# training-step kernels -> gpurun_out/ (summaries are copied into profiles/ by hand)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/tests_r1e.log
tail -3 gpurun_out/tests_r1e.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
tail -c 400 gpurun_out/bench_r1e.err
cut -c1-1500 gpurun_out/bench_r1e.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1e_reference.json 2>/dev/null
timeout 300 python tools/step_timeline.py 2>&1 | tail -24
SMALL="python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline"
timeout 300 $SMALL > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1e.csv $SMALL > gpurun_out/ncu1.log 2>&1
timeout 300 $SMALL > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_train_fwd_tc|k_train_bwd_tc|k_train_mid|k_adam_touched|k_adam_untouched|k_link_ids' --launch-skip 24 -c 12 -f -o gpurun_out/prof_r1e_train $SMALL > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200

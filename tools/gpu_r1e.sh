#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_train_check.py > gpurun_out/tc_check.log 2>&1
echo "tc_check rc=$?"; grep -v " ok$" gpurun_out/tc_check.log | tail -12
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/tests_r1e.log
tail -3 gpurun_out/tests_r1e.log
timeout 300 python tools/step_timeline.py 2>&1 | tail -24 | grep -v "stage_batch"
B="python bench.py --steps 300 --warmup 20 --no-cpu-baseline --eval-users 64"
for cfg in "DCCF_X=1" ; do
  env $cfg timeout 300 $B 2>gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('%-60s'%'$cfg', 'ms_per_step', round(d['ms_per_step'],5), 'b2b', round(d['back_to_back']['ms_per_step'],5), 'e2e', round(d['e2e']['value']), 'loss', d.get('last_loss'))
"
done
tail -3 gpurun_out/bench_err.log

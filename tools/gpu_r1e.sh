#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/tests_r1e.log
tail -2 gpurun_out/tests_r1e.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --eval-users 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value', round(d['value']), 'ms', round(d['ms_per_step'],5), 'b2b', round(d['back_to_back']['ms_per_step'],5), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 'ms/batch', round(d['eval']['ms_per_batch'],4))
for k,v in d['roofline'].get('kernels',{}).items(): print('  ',k, round(v['start_us'],1), round(v['end_us'],1), round(v['us'],1))
"

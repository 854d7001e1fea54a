#!/bin/bash
# one gpurun call (1 GPU): GPU suite, smoke, full bench (ours + reference arm)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/tests_r1e.log
tail -3 gpurun_out/tests_r1e.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
tail -c 400 gpurun_out/bench_r1e.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r1e.json').readline())
print('value', round(d['value']), 'ms', round(d['ms_per_step'],5), 'b2b', round(d['back_to_back']['ms_per_step'],5), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 'eval e2e', round(d['eval']['e2e']['value']), 'launches', d['gpu_launches'], 'cpu', round(d['cpu_baseline']['value']))
for k,v in d['roofline'].get('kernels',{}).items(): print('  ',k, round(v['start_us'],1), round(v['end_us'],1), round(v['us'],1))
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1e_reference.json 2>/dev/null
cut -c1-200 gpurun_out/bench_r1e_reference.json

#!/bin/bash
# Round 2, last 1-GPU call: the bench line exactly as the driver runs it, on the final code, and the step timeline.
mkdir -p gpurun_out
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_n1.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
print('eval', round(d['eval']['value']), 'rank_ms', d['eval']['rank_ms'], 'e2e', round(d['eval']['e2e']['value']))
print('clocks', d.get('clocks'))
P
timeout 100 python tools/step_timeline.py --steps 40 2>&1 | tail -20 > gpurun_out/step_timeline_r2.txt; tail -9 gpurun_out/step_timeline_r2.txt

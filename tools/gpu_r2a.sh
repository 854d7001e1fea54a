#!/bin/bash
# Round 2, GPU call A (1 GPU): suite with the rewritten ranker, ranker micro-bench, bench, A/B build variants,
# ncu captures of the kernels round 1 left unprofiled.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_r2a.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r2a.log
tail -4 gpurun_out/tests_r2a.log
timeout 120 python tools/rank_bench.py --users 1024 16384 2>&1 | tee gpurun_out/rank_bench_r2a.json
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -c 600 gpurun_out/bench_r2a.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 'rank_ms', d['eval']['rank_ms'])
k = d['roofline'].get('kernels', {})
print({n: round(o['us'], 1) for n, o in k.items()})
for kk in ('eval_noise_free', 'eval_projected_noise'):
    o = d.get(kk, {})
    print(kk, {x: o.get(x) for x in ('value', 'ms_per_batch', 'eager_ms_per_batch', 'pass_ms', 'rank_ms', 'parity_ok', 'error') if x in o})
P
for v in s3 p7; do
  DCCF_LIB_VARIANT=$v timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_r2a_ab_$v.json 2>/dev/null
  python - <<P
import json
d = json.loads(open('gpurun_out/bench_r2a_ab_$v.json').read().strip().splitlines()[-1])
k = d['roofline'].get('kernels', {})
print('$v', 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'eval ms/batch', round(d['eval']['ms_per_batch'], 4), {n: round(o['us'], 1) for n, o in k.items()})
P
done
# ncu: the kernels without a capture so far
LEGS="python bench.py --extra-legs-only --eval-users 64"
timeout 300 $LEGS > gpurun_out/plain_legs.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_scores|k_confounder_draw|k_row_scores_tc|k_rank_stream' --launch-skip 8 -c 8 -f -o gpurun_out/prof_r2a_legs $LEGS > gpurun_out/ncu_legs2.log 2>&1
tail -2 gpurun_out/ncu_legs2.log | cut -c1-200
FC="python tools/bench_full_catalogue.py --iters 3 --warmup 1"
timeout 300 $FC > gpurun_out/full_catalogue_r2a.json 2>gpurun_out/fc_r2a.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_full_scores' --launch-skip 2 -c 2 -f -o gpurun_out/prof_r2a_fc $FC > gpurun_out/ncu_fc.log 2>&1
tail -2 gpurun_out/ncu_fc.log | cut -c1-200
cat gpurun_out/full_catalogue_r2a.json | cut -c1-600

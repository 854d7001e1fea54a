#!/bin/bash
# Round 2, GPU call G (1 GPU): full_scores v2 (pre-split item images + TMA ring + 8 epilogue warps), ranker NF=32.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r2g.log
tail -4 gpurun_out/tests_r2g.log
timeout 120 python tools/rank_bench.py --users 1024 16384 2>&1 | tee gpurun_out/rank_bench_r2g.json
timeout 300 python tools/bench_full_catalogue.py --iters 10 --warmup 3 2>gpurun_out/fc_r2g.err | tee gpurun_out/full_catalogue_r2g.json | cut -c1-900
tail -3 gpurun_out/fc_r2g.err
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2g.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2g.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']))
print('eval', round(d['eval']['value']), 'rank_ms', d['eval']['rank_ms'], 'ranker', d['eval'].get('ranker'))
for kk in ('config4_full_catalogue', 'eval_noise_free'):
    print(kk, json.dumps(d.get(kk))[:1500])
P
FC="python tools/bench_full_catalogue.py --iters 3 --warmup 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_full_scores|k_fs_prep_b' --launch-skip 3 -c 3 -f -o gpurun_out/prof_r2g_fc $FC > gpurun_out/ncu_r2g_fc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rank_stream --launch-skip 10 -c 1 -f -o gpurun_out/prof_r2g_rank python tools/rank_bench.py --users 1024 > gpurun_out/ncu_r2g_rank.log 2>&1
ls -la gpurun_out/prof_r2g_*

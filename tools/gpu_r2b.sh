#!/bin/bash
# Round 2, GPU call B (1 GPU): suite, ranker micro-bench, the restructured bench (all legs), step timeline with / without
# the L2 prefetch, ncu captures (launch list of the bench + --set full of the kernels without a round-2 capture).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_r2b.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r2b.log
tail -4 gpurun_out/tests_r2b.log
timeout 120 python tools/rank_bench.py --users 1024 16384 2>&1 | tee gpurun_out/rank_bench_r2b.json
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r2b.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'e2e sync', round(d['e2e']['sync_every_step']['value']))
print('eval', round(d['eval']['value']), 'rank_ms', d['eval']['rank_ms'], 'eval e2e', round(d['eval']['e2e']['value']))
k = d['roofline'].get('kernels', {})
print({n: (round(o['start_us'], 1), round(o['end_us'], 1)) for n, o in k.items()})
print('roof', {x: d['roofline'].get(x) for x in ('kernel', 'achieved', 'frac', 'frac_issued')}, d['roofline'].get('whole_step'))
for kk in ('config3_cds', 'config4_full_catalogue', 'config5_scaled', 'gpu_eager_reference', 'eval_noise_free', 'eval_projected_noise', 'cpu_baseline'):
    o = d.get(kk, {})
    print(kk, json.dumps(o)[:900])
P
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2b_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2b_reference.json
for pf in 1 0; do
  echo "== timeline DCCF_L2_PREFETCH=$pf"; DCCF_L2_PREFETCH=$pf timeout 300 python tools/step_timeline.py --steps 40 2>&1 | tail -18
done
# ncu: launch list of the bench (share of each kernel in the step) and full captures
B="python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline --legs noise_free,projected"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2b.csv $B > gpurun_out/ncu_r2b_1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_train_fwd_tc|k_train_bwd_tc|k_train_mid|k_adam_touched|k_link_ids' --launch-skip 20 -c 10 -f -o gpurun_out/prof_r2b_train python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2b_2.log 2>&1
LEGS="python bench.py --extra-legs-only --eval-users 64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rank_stream' --launch-skip 3 -c 2 -f -o gpurun_out/prof_r2b_rank $LEGS > gpurun_out/ncu_r2b_3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_confounder_draw' --launch-skip 2 -c 2 -f -o gpurun_out/prof_r2b_draw $LEGS > gpurun_out/ncu_r2b_4.log 2>&1
DCCF_EVAL_NOISE=projected timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_row_scores_tc' --launch-skip 6 -c 2 -f -o gpurun_out/prof_r2b_tc64 $LEGS > gpurun_out/ncu_r2b_5.log 2>&1
tail -1 gpurun_out/ncu_r2b_*.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep | tail -8

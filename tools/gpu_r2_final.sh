#!/bin/bash
# Round 2, final 1-GPU call: suite, the bench line as the driver runs it, the reference arm, step timeline, and the ncu
# evidence of the final kernels (launch list of the bench command + --set full captures).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_r2_final.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/tests_r2_final.log
tail -3 gpurun_out/tests_r2_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r2_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_reference.json 2>/dev/null
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
print('eval', round(d['eval']['value']), 'rank_ms', d['eval']['rank_ms'], 'e2e', round(d['eval']['e2e']['value']))
print('roof', {x: d['roofline'].get(x) for x in ('kernel', 'achieved', 'frac', 'frac_issued')}, d['roofline'].get('whole_step'))
for kk in ('config3_cds', 'config4_full_catalogue', 'config5_scaled', 'gpu_eager_reference', 'eval_noise_free', 'eval_projected_noise'):
    print(kk, json.dumps(d.get(kk))[:600])
P
timeout 300 python tools/step_timeline.py --steps 40 2>&1 | tail -20 > gpurun_out/step_timeline_r2.txt; tail -9 gpurun_out/step_timeline_r2.txt
timeout 120 python tools/rank_bench.py --users 1024 16384 > gpurun_out/rank_bench_r2.json 2>&1; cat gpurun_out/rank_bench_r2.json | cut -c1-300
timeout 300 python tools/bench_full_catalogue.py --iters 10 --warmup 3 > gpurun_out/full_catalogue_r2.json 2>/dev/null
B="python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline --legs noise_free,projected"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_r2_1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_train_fwd_tc|k_train_bwd_tc|k_train_mid|k_adam_touched|k_adam_untouched|k_link_ids|k_csr_build' --launch-skip 28 -c 16 -f -o gpurun_out/prof_r2_train python bench.py --steps 6 --warmup 3 --eval-users 64 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2_2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_row_scores_tc|k_backdoor' --launch-skip 4 -c 2 -f -o gpurun_out/prof_r2_eval python bench.py --steps 3 --warmup 3 --eval-users 64 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu_r2_3.log 2>&1
LEGS="python bench.py --extra-legs-only --eval-users 64 --legs noise_free"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gather_scores|k_rank_stream' --launch-skip 6 -c 4 -f -o gpurun_out/prof_r2_gather $LEGS > gpurun_out/ncu_r2_4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rank_stream --launch-skip 10 -c 1 -f -o gpurun_out/prof_r2_rank python tools/rank_bench.py --users 1024 > gpurun_out/ncu_r2_5.log 2>&1
FC="python tools/bench_full_catalogue.py --iters 2 --warmup 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_full_scores|k_fs_prep_b|k_topk_merge' --launch-skip 3 -c 6 -f -o gpurun_out/prof_r2_fc $FC > gpurun_out/ncu_r2_6.log 2>&1
ls -la gpurun_out/prof_r2_*.ncu-rep

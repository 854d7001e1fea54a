#!/bin/bash
# Round 2, GPU call C (2 GPUs): the multi-GPU tests, the data-parallel bench with its parity check and legs, the folded
# exchange synchronisation A/B, the two-stages-per-trip producer variant (pp) A/B.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_r2c.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r2c.log
tail -4 gpurun_out/tests_r2c.log
timeout 120 python tools/rank_bench.py --users 1024 16384 2>&1 | tee gpurun_out/rank_bench_r2c.json
for v in "" pp; do
  echo "== timeline variant '${v}'"; DCCF_LIB_VARIANT=$v timeout 300 python tools/step_timeline.py --steps 40 2>&1 | tail -9
done
DCCF_LIB_VARIANT=pp timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "training or fused or graph or resident or config0" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_r2c_dp2.json 2> gpurun_out/bench_r2c_dp2.err
echo "dp2 rc=$?"; tail -c 1200 gpurun_out/bench_r2c_dp2.err
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/bench_r2c_dp2.json').read().strip().splitlines()[-1])
    print('dp2 value', d.get('value'), 'ms', d.get('ms_per_step'), 'b2b', d.get('back_to_back', {}).get('ms_per_step'), 'e2e', d.get('e2e', {}).get('value'))
    print('parity', json.dumps(d.get('dp_parity'))[:900])
    k = d['roofline'].get('kernels', {})
    print({n: (round(o['start_us'], 1), round(o['end_us'], 1)) for n, o in k.items()})
    for kk in ('config3_cds', 'config4_full_catalogue', 'config5_scaled'):
        print(kk, json.dumps(d.get(kk))[:700])
    print('eval', d['eval']['value'], d['eval']['rank_ms'])
except Exception as e:
    print('parse failed', e)
P
for fold in 1 0; do
  DCCF_DP_FOLD=$fold timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-extra-legs > gpurun_out/bench_r2c_dp2_fold$fold.json 2>/dev/null
  python - <<P
import json
try:
    d = json.loads(open('gpurun_out/bench_r2c_dp2_fold$fold.json').read().strip().splitlines()[-1])
    k = d['roofline'].get('kernels', {})
    print('fold=$fold', 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'parity', d.get('dp_parity_ok'), {n: (round(o['start_us'], 1), round(o['end_us'], 1)) for n, o in k.items()})
except Exception as e:
    print('fold=$fold parse failed', e)
P
done
DCCF_LIB_VARIANT=pp timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-extra-legs > gpurun_out/bench_r2c_dp2_pp.json 2>/dev/null
python - <<'P'
import json
try:
    d = json.loads(open('gpurun_out/bench_r2c_dp2_pp.json').read().strip().splitlines()[-1])
    print('pp dp2', 'ms/step', round(d['ms_per_step'], 5), 'parity', d.get('dp_parity_ok'))
except Exception as e:
    print('pp parse failed', e)
P

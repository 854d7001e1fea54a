#!/bin/bash
# Round 2, final 2-/4-GPU call: the full bench line at 2 ranks (every leg), the training / evaluation line at 4 ranks.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
EXTRA=${2:-}
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 $EXTRA > gpurun_out/bench_r2_dp$N.json 2> gpurun_out/bench_r2_dp$N.err; echo rc=$?; tail -c 300 gpurun_out/bench_r2_dp$N.err
python - gpurun_out/bench_r2_dp$N.json <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['roofline'].get('kernels', {})
print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'parity', d.get('dp_parity_ok'), 'eval', round(d['eval']['value']))
for n, o in sorted(k.items(), key=lambda kv: kv[1]['start_us']):
    print('    %-32s %6.1f -> %6.1f' % (n, o['start_us'], o['end_us']))
for kk in ('config3_cds', 'config4_full_catalogue', 'config5_scaled'):
    if kk in d: print(kk, json.dumps(d.get(kk))[:500])
P

#!/bin/bash
# Round 2, GPU call X (1 GPU): the final library on the GPU suite, then the programmatic chain fwd -> mid -> bwd
# (DCCF_PDL_CHAIN=1) on the training tests and the bench line beside the default.
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
DCCF_PDL_CHAIN=1 timeout 90 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_training or graph_replay or tensor_core_training or resident_epoch or training_state" 2>&1 | tail -2
for f in 1 0; do
  DCCF_PDL_CHAIN=$f timeout 100 python bench.py --steps 100 --warmup 5 --no-extra-legs --no-cpu-baseline --eval-users 64 > gpurun_out/bench_r2x_pdl$f.json 2>/dev/null
  python - gpurun_out/bench_r2x_pdl$f.json <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d['roofline'].get('kernels', {})
print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']))
print('   ', ' | '.join('%s %.1f-%.1f' % (n.replace('k_', ''), o['start_us'], o['end_us']) for n, o in sorted(k.items(), key=lambda kv: kv[1]['start_us'])))
P
done

#!/bin/bash
# Round 2, GPU call W (2 GPUs): CSR ranges after the middle kernel on a fourth stream (DCCF_CSR_PLACE=after_mid, the default
# above 2 ranks) forced at 2 ranks, against the default there (beside).
mkdir -p gpurun_out
DCCF_CSR_PLACE=after_mid timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
show() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d['roofline'].get('kernels', {})
    print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'parity', d.get('dp_parity_ok'))
    for n, o in sorted(k.items(), key=lambda kv: kv[1]['start_us']):
        print('    %-32s %6.1f -> %6.1f' % (n, o['start_us'], o['end_us']))
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
P
}
for f in after_mid beside; do
  echo "== DCCF_CSR_PLACE=$f"
  DCCF_CSR_PLACE=$f timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-extra-legs --eval-users 256 > gpurun_out/bench_r2w_dp2_$f.json 2> gpurun_out/bench_r2w_dp2_$f.err; echo rc=$?; tail -c 300 gpurun_out/bench_r2w_dp2_$f.err; show gpurun_out/bench_r2w_dp2_$f.json
done

#!/bin/bash
# Round 2, GPU call O (2 GPUs): k_link_ids beside the forward (DCCF_LINK_BESIDE) x width of the side sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests_r2o.log
tail -3 gpurun_out/tests_r2o.log
for v in "0 128" "1 128" "1 160" "1 192"; do
  set -- $v
  echo "== DCCF_LINK_BESIDE=$1 DCCF_SIDE_THREADS=$2"
  DCCF_LINK_BESIDE=$1 DCCF_SIDE_THREADS=$2 timeout 300 python tools/step_timeline.py --steps 40 2>&1 | tail -18
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
show() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d['roofline'].get('kernels', {})
    print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5), 'e2e', round(d['e2e']['value']), 'parity', d.get('dp_parity_ok'), 'eval', round(d['eval']['value']), 'rank_ms', round(d['eval']['rank_ms'], 4))
    for n, o in sorted(k.items(), key=lambda kv: kv[1]['start_us']):
        print('    %-32s %6.1f -> %6.1f' % (n, o['start_us'], o['end_us']))
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
P
}
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-extra-legs > gpurun_out/bench_r2o_dp2.json 2> gpurun_out/bench_r2o_dp2.err; echo rc=$?; tail -c 300 gpurun_out/bench_r2o_dp2.err; show gpurun_out/bench_r2o_dp2.json
timeout 600 python bench.py --steps 100 --warmup 5 --no-extra-legs > gpurun_out/bench_r2o_1.json 2>/dev/null; show gpurun_out/bench_r2o_1.json
DCCF_SIDE_THREADS=160 timeout 600 python bench.py --steps 100 --warmup 5 --no-extra-legs > gpurun_out/bench_r2o_1b.json 2>/dev/null; show gpurun_out/bench_r2o_1b.json

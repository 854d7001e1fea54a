#!/bin/bash
# one gpurun --gpus N call: data-parallel check (replica identity, equality with the single-GPU step) and the
# N-GPU bench line with the step timeline of rank 0
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py > gpurun_out/dp_check_$N.log 2>&1
grep -i "error\|assert\|DP CHECK\|Traceback" gpurun_out/dp_check_$N.log | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 10 --eval-users 256 > gpurun_out/bench_r1e_dp$N.json 2> gpurun_out/bench_r1e_dp$N.err
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_r1e_dp$N.json') if l.startswith('{')][-1])
print('dp$N value', round(d['value']), 'ms', round(d['ms_per_step'],5), 'b2b', round(d['back_to_back']['ms_per_step'],5), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']))
for k,v in d['roofline'].get('kernels',{}).items(): print('  ',k, round(v['start_us'],1), round(v['end_us'],1), round(v['us'],1))
"

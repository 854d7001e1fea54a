#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "row_sharded or tensor_core_training or graph_replay or resident" 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dp_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 10 --eval-users 256 > gpurun_out/bench_r1e_dp2.json 2> gpurun_out/bench_r1e_dp2.err
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_r1e_dp2.err | tail -12
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1e_dp2.json') if l.startswith('{')][-1])
print('dp2 value', round(d['value']), 'ms', round(d['ms_per_step'],5), 'b2b', round(d['back_to_back']['ms_per_step'],5), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']))
for k,v in d['roofline'].get('kernels',{}).items(): print('  ',k, round(v['start_us'],1), round(v['end_us'],1), round(v['us'],1))
"

#!/bin/bash
# Round 2, N GPUs (gpurun --gpus 2 ...): first hardware run of src/main.py under torchrun (data-parallel training through
# the reference CLI: --batch_size per rank, sharded evaluation, rank 0 owns the files; ends with a replica checksum).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2_dp_cli.sh 2'
N=${1:-2}
mkdir -p gpurun_out /tmp/dpcli
python -m dccf_b200.synth --path /tmp/dpcli/datasets/ --dataset toy --preset tiny > /dev/null
cd src
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    main.py --rank 1 --model_name DCCF --optimizer Adam --lr 0.001 --dataset toy --path /tmp/dpcli/datasets/ \
    --metric ndcg@5,recall@5,precision@5 --epoch 3 --test_neg_n 100 --log_file /tmp/dpcli/log.txt \
    --result_file /tmp/dpcli/result.npy --model_path /tmp/dpcli/model/m.pt > ../gpurun_out/dp_cli_$N.log 2>&1
echo "rc=$?"
cd ..
grep -E "data parallel|Epoch|Test (Before|After)|Best Iter|diverged|Error|error" gpurun_out/dp_cli_$N.log | tail -20
# the scaled configuration (row-sharded users + on-the-fly IPS-MF exposure + data parallel): equality with one GPU at a
# small shape, then ms / step at 10 M users x 1 M items split over the N ranks
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    tools/scaled_check.py > gpurun_out/scaled_check_$N.log 2>&1
echo "scaled rc=$?"
grep -E "scaled|SCALED|Error|error" gpurun_out/scaled_check_$N.log | tail -8
# A/B of the data-parallel push: arrival flags published by one release store per lane (variant pf) against the sequence of
# `world` release stores by one thread (default); dp_check first (correctness), then the bench at N ranks
DCCF_LIB_VARIANT=pf DCCF_BUILD_DEFS="-DDCCF_DP_PARALLEL_FLAGS" python -m dccf_b200.build > /dev/null
for v in "" pf; do
  DCCF_LIB_VARIANT=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29515 tools/dp_check.py 2>&1 | grep -E "DP CHECK|Error|assert" | tail -2
  DCCF_LIB_VARIANT=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29516 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_r2a_dp${N}_${v:-base}.json 2>/dev/null
  python - <<P
import json
d = json.loads(open('gpurun_out/bench_r2a_dp${N}_${v:-base}.json').read().strip().splitlines()[-1])
print('dp$N ${v:-base}', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 5), 'b2b', round(d['back_to_back']['ms_per_step'], 5))
P
done

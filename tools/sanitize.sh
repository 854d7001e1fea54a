#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (run on a B200 box, one GPU; slow: small shapes only).
#   bash tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
# memcheck / initcheck cover every kernel the check script and the small parity tests launch; racecheck and
# synccheck look at the shared-memory pipelines (mbarrier rings of the tensor-core kernels, the middle kernel's
# phases).  Output under gpurun_out/sanitize_<tool>.log.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 3 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "tensor_core_training_step_equals_simt_step and 50-70 or row_sharded or forward_backward_vs_oracle and 20-30 or adam_sweep or ranker_topk" \
    > gpurun_out/sanitize_"$TOOL".log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
tail -5 gpurun_out/sanitize_"$TOOL".log
# the kernels of round 1 that were written without hardware (gather scorer, device confounder draw): small cases only
timeout 900 compute-sanitizer --tool "$TOOL" --error-exitcode 3 \
    python tests/test_gpu_zz_gather.py oracle_ragged_many_slots oracle_no_confounders out_of_range_ids \
    device_confounder_draw_equals_torch_randint > gpurun_out/sanitize_"$TOOL"_blind.log 2>&1
echo "compute-sanitizer $TOOL (blind kernels) rc=$?"
grep -E "GATHER_(OK|FAIL)|ERROR SUMMARY" gpurun_out/sanitize_"$TOOL"_blind.log | tail -8

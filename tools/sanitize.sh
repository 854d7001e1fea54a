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

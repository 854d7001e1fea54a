#!/bin/bash
mkdir -p gpurun_out
for cfg in "DCCF_X=1" "DCCF_TC_KSPLITS=6 DCCF_TC_BWD_SPLITS=42"; do
  echo "=== $cfg"
  env $cfg timeout 300 python tools/step_timeline.py 2>&1 | tail -22
done

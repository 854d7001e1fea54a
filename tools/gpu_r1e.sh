#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py 2>&1 | tail -24 | grep -v "stage_batch"

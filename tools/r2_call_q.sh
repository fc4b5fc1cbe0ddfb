#!/bin/bash
# round 2, GPU call Q: timeline of one replayed cfg2 step under the row-wise pipeline settings
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py timeline > gpurun_out/r2q_timeline.log 2>&1; echo "timeline rc=$?"; tail -n 30 gpurun_out/r2q_timeline.log

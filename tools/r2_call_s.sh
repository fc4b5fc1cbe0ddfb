#!/bin/bash
# round 2, GPU call S: explicit launch priorities (recorded in the captured graph) x row-wise pipeline settings
mkdir -p gpurun_out
export ROWP_COMBOS="0:2048:1:4:1,1:2048:1:1:0,1:2048:1:4:1,1:2048:1:8:1,1:2048:2:4:1,1:2048:2:8:1,1:4096:2:8:1"
for pr in 1 0; do
MOGP_LAUNCH_PRIO=$pr DIAG_CFGS=cfg1,cfg2,cfg4 timeout 600 python tools/gpu_diag.py rowp > gpurun_out/r2s_rowp_prio$pr.log 2>&1; echo "rowp prio=$pr rc=$?"; grep "step" gpurun_out/r2s_rowp_prio$pr.log | tail -n 30
MOGP_LAUNCH_PRIO=$pr timeout 300 python tools/gpu_diag.py timeline > gpurun_out/r2s_timeline_prio$pr.log 2>&1; echo "timeline prio=$pr rc=$?"; grep "^\[timeline" gpurun_out/r2s_timeline_prio$pr.log
done
DIAG_CFGS=cfg3 ROWP_COMBOS="0:2048:1:4:1" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep step
MOGP_LAUNCH_PRIO=0 DIAG_CFGS=cfg3 ROWP_COMBOS="0:2048:1:4:1" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep step

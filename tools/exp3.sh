#!/bin/bash
# Experiment 3: two-scalar-warp panel step (variant 4) against variant 2, with the pipelined inverse.
mkdir -p gpurun_out
timeout 100 python -u tools/gpu_diag.py peak > gpurun_out/exp3_peak.log 2>&1; echo "peak rc=$?"
EXP_COMBOS="4:0,4:1,2:1" DIAG_CFGS=cfg2,cfg4,cfg3 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp3_v4.log 2>&1; echo "exp v4 rc=$?"
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp3_panel.log 2>&1; echo "panel rc=$?"
grep -h "issue\|VERDICT\|potrf n=\|step \|panel\|chain" gpurun_out/exp3_peak.log gpurun_out/exp3_v4.log gpurun_out/exp3_panel.log | grep -v "relerr(L)"
tail -n 3 gpurun_out/exp3_v4.log

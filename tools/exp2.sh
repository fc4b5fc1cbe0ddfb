#!/bin/bash
# Experiment 2: decoupled warp-specialised panel step (variant 3) and the multi-stream pipelined inverse.
mkdir -p gpurun_out
EXP_COMBOS="2:0,2:1" DIAG_CFGS=cfg2,cfg4 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp2_v2.log 2>&1; echo "exp v2 rc=$?"
EXP_COMBOS="3:0,3:1" DIAG_CFGS=cfg2,cfg4,cfg3 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp2_v3.log 2>&1; echo "exp v3 rc=$?"
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp2_panel.log 2>&1; echo "panel rc=$?"
grep -h "VERDICT\|potrf n=\|step \|panel\|chain" gpurun_out/exp2_v*.log gpurun_out/exp2_panel.log | grep -v "relerr(L)"
tail -n 5 gpurun_out/exp2_v3.log

#!/bin/bash
# Experiment 5: tensor warps with compile-time tile masks (no predicated DMMA), variants 2 and 4.
mkdir -p gpurun_out
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp5_panel.log 2>&1; echo "panel rc=$?"
EXP_COMBOS="2:1,4:1,1:1" DIAG_CFGS=cfg2,cfg4,cfg3 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp5.log 2>&1; echo "exp rc=$?"
grep -h "VERDICT\|potrf n=\|step \|panel\|chain\|pivot warp\|---" gpurun_out/exp5_panel.log gpurun_out/exp5.log | grep -v "relerr(L)"

#!/bin/bash
# Experiment 4: cost of the block-scope fences before the named-barrier arrivals; phase stamps of the pivot warp.
mkdir -p gpurun_out
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp4_panel.log 2>&1; echo "panel rc=$?"
EXP_NOFENCE=1 EXP_COMBOS="2:1,4:1" DIAG_CFGS=cfg2,cfg4 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp4_nofence.log 2>&1; echo "exp rc=$?"
grep -h "VERDICT\|potrf n=\|step \|panel\|chain\|pivot warp\|---" gpurun_out/exp4_panel.log gpurun_out/exp4_nofence.log | grep -v "relerr(L)"

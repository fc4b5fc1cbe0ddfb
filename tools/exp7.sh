#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp7_panel.log 2>&1; echo "panel rc=$?"
EXP_NOFENCE=${NOFENCE:-0} EXP_COMBOS="${COMBOS:-4:1,2:1}" DIAG_CFGS=cfg2,cfg4 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp7.log 2>&1; echo "exp rc=$?"
grep -h "VERDICT\|potrf n=\|step \|panel\|chain\|pivot warp\|tensor warp\|---" gpurun_out/exp7_panel.log gpurun_out/exp7.log | grep -v "relerr(L)"

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -u tools/gpu_diag.py spans > gpurun_out/exp13_spans.log 2>&1; echo "spans rc=$?"; grep "spans\|pdl" gpurun_out/exp13_spans.log
EXP_PDL=1 EXP_COMBOS="2:1" DIAG_CFGS=cfg2,cfg4,cfg3 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp13_pdl1.log 2>&1; echo "exp rc=$?"
EXP_PDL=0 EXP_COMBOS="2:1" DIAG_CFGS=cfg2,cfg4 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp13_pdl0.log 2>&1; echo "exp rc=$?"
grep -h "pdl\|VERDICT\|potrf n=\|step " gpurun_out/exp13_pdl1.log gpurun_out/exp13_pdl0.log | grep -v "relerr(L)"
tail -3 gpurun_out/exp13_pdl1.log

#!/bin/bash
# round 2, GPU call X: recursive scheme with the trailing products overlapped with the second half's recursion (on / off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "recursive" > gpurun_out/r2x_pytest_rchol.log 2>&1; echo "pytest rchol rc=$?"; tail -n 5 gpurun_out/r2x_pytest_rchol.log
for ov in 1 0; do
MOGP_RCHOL_OVERLAP=$ov DIAG_CFGS=cfg4,cfg3 EXP_COMBOS="2:1" timeout 600 python tools/gpu_diag.py exp 2>&1 | grep "potrf n=8192\|potrf n=4096\|step" | sed "s/^/overlap=$ov /"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg3 or cfg4" > gpurun_out/r2x_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -n 3 gpurun_out/r2x_pytest_parity.log

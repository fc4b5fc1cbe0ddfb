#!/bin/bash
# round 2, GPU call W: recursive Cholesky + inverse on the int8 pipe (N >= 4096): numerics, step and potrf times on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "recursive" > gpurun_out/r2w_pytest_rchol.log 2>&1; echo "pytest rchol rc=$?"; tail -n 25 gpurun_out/r2w_pytest_rchol.log
for rc in 1 0; do
MOGP_RCHOL=$rc DIAG_CFGS=cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 600 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/rchol=$rc /"
MOGP_RCHOL=$rc DIAG_CFGS=cfg4,cfg3 EXP_COMBOS="2:1" timeout 600 python tools/gpu_diag.py exp 2>&1 | grep "potrf n=\|step\|VERDICT" | sed "s/^/rchol=$rc /"
done

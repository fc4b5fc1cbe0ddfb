#!/bin/bash
# round 2, GPU call T: row-wise pipeline with / without the progressive K^-1, S = 2, N = 4096; then the whole GPU suite on the new defaults
mkdir -p gpurun_out
export ROWP_COMBOS="0:2048:1:1:0,1:2048:1:1:0,1:2048:1:2:0,1:2048:1:2:1,1:4096:1:1:0"
for kv in 1 0; do
MOGP_ROWPIPE_KINV=$kv DIAG_CFGS=cfg1,cfg2,cfg4 timeout 600 python tools/gpu_diag.py rowp > gpurun_out/r2t_rowp_kinv$kv.log 2>&1; echo "rowp kinv=$kv rc=$?"; grep "step" gpurun_out/r2t_rowp_kinv$kv.log | tail -n 30
done
MOGP_ROWPIPE_KINV=0 ROWP_COMBOS="1:2048:1:1:0" timeout 300 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2t_pytest_gpu.log

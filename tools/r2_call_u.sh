#!/bin/bash
# round 2, GPU call U: K^-1 accumulation in halving chunks (mode 2) against modes 0 / 1; knob tests
mkdir -p gpurun_out
export ROWP_COMBOS="0:2048:1:1:0,1:4096:1:1:0"
for kv in "0 4" "2 4" "2 8" "2 2" "1 4"; do set -- $kv
MOGP_ROWPIPE_KINV=$1 MOGP_ROWPIPE_WMIN=$2 DIAG_CFGS=cfg1,cfg2,cfg4 timeout 600 python tools/gpu_diag.py rowp > gpurun_out/r2u_rowp_kinv$1_$2.log 2>&1; echo "rowp kinv=$1 wmin=$2 rc=$?"; grep "step" gpurun_out/r2u_rowp_kinv$1_$2.log | tail -n 30
MOGP_ROWPIPE_KINV=$1 MOGP_ROWPIPE_WMIN=$2 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline"
done
timeout 1500 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x > gpurun_out/r2u_pytest_knobs.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2u_pytest_knobs.log

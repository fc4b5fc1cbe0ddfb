#!/bin/bash
# round 2, GPU call AB: single-shot rank-64 GEMM kernel (one memory round trip per tile)
mkdir -p gpurun_out
GEMMK_SHAPES="2048x2048x64,1024x1024x64,4096x4096x64,8192x8192x64,512x2048x64" timeout 300 python tools/gpu_diag.py gemmk > gpurun_out/r2ab_gemmk.log 2>&1; echo "gemmk rc=$?"; grep "gemm NT" gpurun_out/r2ab_gemmk.log
for k in 1 0; do
MOGP_GEMM_K64=$k DIAG_CFGS=cfg1,cfg2,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/k64=$k /"
MOGP_GEMM_K64=$k ROWP_COMBOS="1:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline" | sed "s/^/k64=$k /"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2ab_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -n 3 gpurun_out/r2ab_pytest_parity.log

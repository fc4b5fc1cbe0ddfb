#!/bin/bash
# round 2, GPU call AA: GEMM epilogue with the old C values of a fragment row loaded together (rank-64 updates)
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py gemmk > gpurun_out/r2aa_gemmk.log 2>&1; echo "gemmk rc=$?"; grep "gemm NT" gpurun_out/r2aa_gemmk.log
DIAG_CFGS=cfg1,cfg2,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step"
ROWP_COMBOS="1:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2aa_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -n 3 gpurun_out/r2aa_pytest_parity.log

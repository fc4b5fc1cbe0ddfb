#!/bin/bash
# round 2, GPU call AD: recursive scheme with leaves other than 2048 rows (Np / 2^k), against the blocked sweep
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py rsizes > gpurun_out/r2ad_rsizes.log 2>&1; echo "rsizes rc=$?"; grep "rsizes" gpurun_out/r2ad_rsizes.log
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "recursive" > gpurun_out/r2ad_pytest_rchol.log 2>&1; echo "pytest rchol rc=$?"; tail -n 4 gpurun_out/r2ad_pytest_rchol.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2ad_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -n 3 gpurun_out/r2ad_pytest_parity.log
DIAG_CFGS=cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step"

#!/bin/bash
# round 2, GPU call AE: padding policy (a few more rows to make large sizes eligible for the recursive scheme) + the whole GPU suite
mkdir -p gpurun_out
RS_SIZES="4224,5000,6000,7000,7500,8000" timeout 600 python tools/gpu_diag.py rsizes > gpurun_out/r2ae_rsizes.log 2>&1; echo "rsizes rc=$?"; grep "rsizes N" gpurun_out/r2ae_rsizes.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ae_pytest_gpu.log

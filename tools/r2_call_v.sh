#!/bin/bash
# round 2, GPU call V: MOHSM on the product path (K, K_diag, LML, gradients, predictions, drop-in under mogptk.MOHSM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "further_kernel or mohsm" > gpurun_out/r2v_pytest_mohsm.log 2>&1; echo "pytest mohsm rc=$?"; tail -n 30 gpurun_out/r2v_pytest_mohsm.log
timeout 900 python -m pytest tests/test_gpu_reference_dropin.py -m gpu -q -x -k "csm_and_sm_lmc" > gpurun_out/r2v_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -n 30 gpurun_out/r2v_pytest_dropin.log

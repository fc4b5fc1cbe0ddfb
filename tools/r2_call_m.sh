#!/bin/bash
# round 2, GPU call M: initialisers through the engine
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_dropin.py -m gpu -q -x -k "bnse or init_parameters" > gpurun_out/r2m_pytest_init.log 2>&1; echo "pytest init rc=$?"; tail -n 30 gpurun_out/r2m_pytest_init.log

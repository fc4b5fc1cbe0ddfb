#!/bin/bash
# round 2, GPU call N: GEMM three ways, fit_adam scratch caching, bench with the per-step device-Adam record
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py gemm3 > gpurun_out/r2n_gemm3.log 2>&1; echo "gemm3 rc=$?"; tail -n 12 gpurun_out/r2n_gemm3.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_dropin.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2n_pytest.log
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); e=d['e2e']; print(d['value'], e['value'], e['fused_optimizer']['value'], e['per_step_device_adam']['value'], e['c_abi_host_call']['value'], e['device_resident_training']['value'])"; tail -n 3 gpurun_out/r2n_bench.err

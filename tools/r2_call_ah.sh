#!/bin/bash
# round 2, GPU call AH: z = L^-1 y along the row-wise pipeline, early loss from inside potrf_padded
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity.py tests/test_gpu_reference_dropin.py -m gpu -q -x > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ah_pytest.log
timeout 600 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x > gpurun_out/r2ah_pytest_knobs.log 2>&1; echo "pytest knobs rc=$?"; tail -n 3 gpurun_out/r2ah_pytest_knobs.log
timeout 600 python bench.py --steps 200 --no-extras --no-cpu-baseline > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2ah_bench.json')); e=d['e2e']; print('value', round(d['value'],1), 'e2e', round(e['value'],1), 'fused', round(e['fused_optimizer']['value'],1), 'dev_adam', round(e['per_step_device_adam']['value'],1), 'cabi', round(e['c_abi_host_call']['value'],1), 'resident', round(e['device_resident_training']['value'],1))"
ROWP_COMBOS="1:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline"

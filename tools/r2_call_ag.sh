#!/bin/bash
# round 2, GPU call AG: early loss (loss() returns when [lml, info] are known; the gradient completes in stream order)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_dropin.py -m gpu -q -x > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2ag_pytest.log
for el in 1 0; do
MOGP_EARLY_LOSS=$el timeout 600 python bench.py --steps 200 --no-extras --no-cpu-baseline > gpurun_out/r2ag_bench_early$el.json 2> gpurun_out/r2ag_bench_early$el.err; python -c "
import json; d=json.load(open('gpurun_out/r2ag_bench_early$el.json')); e=d['e2e']; print('early=$el value', round(d['value'],1), 'e2e', round(e['value'],1), 'fused', round(e['fused_optimizer']['value'],1), 'dev_adam', round(e['per_step_device_adam']['value'],1), 'cabi', round(e['c_abi_host_call']['value'],1), 'resident', round(e['device_resident_training']['value'],1))"
done

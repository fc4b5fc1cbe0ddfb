#!/bin/bash
# round 2, GPU call E: int8 trailing updates of the Cholesky, full suite, cfg3/cfg4 bench with / without them
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -k "int8" > gpurun_out/r2e_pytest_i8.log 2>&1; echo "pytest i8 rc=$?"; tail -n 12 gpurun_out/r2e_pytest_i8.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2e_pytest.log
for cfg in cfg3 cfg4; do
for pm in 8192 4096 0; do
MOGP_I8_POTRF_MIN=$pm timeout 300 python bench.py --config $cfg --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_${cfg}_p$pm.json 2> gpurun_out/r2e_bench_${cfg}_p$pm.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_${cfg}_p$pm.json')); print('$cfg potrf_min=$pm', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done; done

#!/bin/bash
# round 2, GPU call H: three-level Cholesky with int8 rank-1024 super-panel updates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -k "int8" > gpurun_out/r2h_pytest_i8.log 2>&1; echo "pytest i8 rc=$?"; tail -n 8 gpurun_out/r2h_pytest_i8.log
for cfg in cfg3 cfg4; do
for pm in 0 4096; do
MOGP_I8_POTRF_MIN=$pm timeout 300 python bench.py --config $cfg --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench_${cfg}_p$pm.json 2> gpurun_out/r2h_bench_${cfg}_p$pm.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_${cfg}_p$pm.json')); print('$cfg potrf_min=$pm', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done; done
MOGP_I8_POTRF_MIN=4096 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_dropin.py -m gpu -q -k "cfg3 or cfg4" > gpurun_out/r2h_pytest_big.log 2>&1; echo "pytest big (three-level on) rc=$?"; tail -n 5 gpurun_out/r2h_pytest_big.log

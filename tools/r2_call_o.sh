#!/bin/bash
# round 2, GPU call O: int8 K^-1 at N = 2048?
mkdir -p gpurun_out
for mn in 4096 2048; do
for rep in 1 2; do
MOGP_I8_MIN_NP=$mn timeout 300 python bench.py --config cfg2 --steps 200 --no-extras --no-cpu-baseline > gpurun_out/r2o_bench_cfg2_mn${mn}_$rep.json 2> gpurun_out/r2o_bench_cfg2_mn${mn}_$rep.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_cfg2_mn${mn}_$rep.json')); print('cfg2 i8_min_np=$mn', d['value'], d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['device_resident_training']['value'])"
done; done
MOGP_I8_MIN_NP=2048 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest (i8 from 2048) rc=$?"; tail -n 3 gpurun_out/r2o_pytest.log

#!/bin/bash
# round 2, GPU call D: int8 triangular-inverse levels, new kernel families (once wired), cfg3 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -k "int8" > gpurun_out/r2d_pytest_i8.log 2>&1; echo "pytest i8 rc=$?"; tail -n 12 gpurun_out/r2d_pytest_i8.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2d_bench_cfg3.json 2> gpurun_out/r2d_bench_cfg3.err; echo "bench cfg3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_cfg3.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
